// datastep.cu — the step in front of the path: liblcg's binary COO fixtures -> CSR on the
// device, without cuSPARSE.  The reference's samples read the file on the host (sample8.cu:30-64), copy the triplets
// over and call cusparseXcoo2csr on the row-sorted indices (sample8.cu:169); lcgb200_coo2csr is that call.
#include "common.cuh"
#include "../../include/lcgb200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <numeric>
#include <algorithm>

namespace lcgb200 {

// rows[] ascending.  row_ptr[r] = first k with rows[k] >= r; row_ptr[n] = nnz.  One thread per entry boundary.
__global__ void k_coo2csr(const int* __restrict__ rows, int nnz, int n, int* __restrict__ row_ptr)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > nnz) return;
	const int prev = (k == 0) ? -1 : rows[k - 1];
	const int cur = (k == nnz) ? n : rows[k];
	for (int r = prev + 1; r <= cur; r++) row_ptr[r] = k;   // empty rows in between get the same offset
}

// 1 in *flag when rows[] is not ascending or leaves [0, n)
__global__ void k_check_sorted(const int* __restrict__ rows, int nnz, int n, int* flag)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nnz) return;
	const int r = rows[k];
	if (r < 0 || r >= n || (k > 0 && rows[k - 1] > r)) *flag = 1;
}

}  // namespace lcgb200

using namespace lcgb200;

extern "C" {

int lcgb200_coo2csr(const int* rows_dev, int nnz, int n, int* row_ptr_dev, void* stream)
{
	if (!rows_dev || !row_ptr_dev) return LCGB200_INVALID_POINTER;
	if (n <= 0 || nnz < 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	cudaStream_t s = (cudaStream_t)stream;
	int* flag = nullptr;
	if (cudaMalloc((void**)&flag, sizeof(int)) != cudaSuccess) return LCGB200_UNKNOWN_ERROR;
	cudaMemsetAsync(flag, 0, sizeof(int), s);
	if (nnz > 0) k_check_sorted<<<(nnz + 255) / 256, 256, 0, s>>>(rows_dev, nnz, n, flag);
	int bad = 0;
	cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
	cudaStreamSynchronize(s);
	cudaFree(flag);
	if (bad) { set_error_msg("lcgb200_coo2csr needs row indices in [0,n) in ascending order (like cusparseXcoo2csr)"); return LCGB200_SIZE_NOT_MATCH; }
	k_coo2csr<<<(nnz + 1 + 255) / 256, 256, 0, s>>>(rows_dev, nnz, n, row_ptr_dev);
	return cudaGetLastError() == cudaSuccess ? 0 : LCGB200_UNKNOWN_ERROR;
}

// Reads a reference fixture (data/README:1-10; the [d] block the README mentions is absent from the files):
//   case_*_A : int32 N | int32 nz | nz x { int32 row, int32 col, value } | N x value   (value: double or 2 doubles)
// Outputs are malloc'ed host arrays the caller frees with lcgb200_free_host; triplets are returned split and,
// if the file is not row-sorted, stably sorted by row.
int lcgb200_read_case(const char* path_A, int value_type, int* n_out, int* nz_out, int** rows_out, int** cols_out, void** vals_out, void** rhs_out)
{
	if (!path_A || !n_out || !nz_out || !rows_out || !cols_out || !vals_out || !rhs_out) return LCGB200_INVALID_POINTER;
	const size_t vs = value_type == LCGB200_REAL ? sizeof(double) : 2 * sizeof(double);
	FILE* f = std::fopen(path_A, "rb");
	if (!f) { set_error_msg("cannot open the case file"); return LCGB200_INVALID_POINTER; }
	int n = 0, nz = 0;
	bool ok = std::fread(&n, 4, 1, f) == 1 && std::fread(&nz, 4, 1, f) == 1 && n > 0 && nz >= 0;
	int* rows = nullptr; int* cols = nullptr; char* vals = nullptr; char* rhs = nullptr;
	if (ok)
	{
		rows = (int*)std::malloc(sizeof(int) * (size_t)std::max(nz, 1)); cols = (int*)std::malloc(sizeof(int) * (size_t)std::max(nz, 1));
		vals = (char*)std::malloc(vs * (size_t)std::max(nz, 1)); rhs = (char*)std::malloc(vs * (size_t)n);
		ok = rows && cols && vals && rhs;
		const size_t rec = 8 + vs;
		std::vector<char> buf((size_t)std::max(nz, 1) * rec);
		ok = ok && std::fread(buf.data(), rec, (size_t)nz, f) == (size_t)nz;
		bool sorted = true;
		for (int k = 0; ok && k < nz; k++)
		{
			std::memcpy(&rows[k], &buf[(size_t)k * rec], 4); std::memcpy(&cols[k], &buf[(size_t)k * rec + 4], 4);
			std::memcpy(vals + (size_t)k * vs, &buf[(size_t)k * rec + 8], vs);
			if (rows[k] < 0 || rows[k] >= n || cols[k] < 0 || cols[k] >= n) ok = false;
			if (k > 0 && rows[k - 1] > rows[k]) sorted = false;
		}
		ok = ok && std::fread(rhs, vs, (size_t)n, f) == (size_t)n;
		if (ok && !sorted)
		{
			std::vector<int> order((size_t)nz); std::iota(order.begin(), order.end(), 0);
			std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rows[a] < rows[b]; });
			std::vector<int> r2((size_t)nz), c2((size_t)nz); std::vector<char> v2((size_t)nz * vs);
			for (int k = 0; k < nz; k++) { r2[(size_t)k] = rows[order[(size_t)k]]; c2[(size_t)k] = cols[order[(size_t)k]]; std::memcpy(&v2[(size_t)k * vs], vals + (size_t)order[(size_t)k] * vs, vs); }
			std::memcpy(rows, r2.data(), sizeof(int) * (size_t)nz); std::memcpy(cols, c2.data(), sizeof(int) * (size_t)nz); std::memcpy(vals, v2.data(), vs * (size_t)nz);
		}
	}
	std::fclose(f);
	if (!ok) { std::free(rows); std::free(cols); std::free(vals); std::free(rhs); set_error_msg("malformed case file"); return LCGB200_SIZE_NOT_MATCH; }
	*n_out = n; *nz_out = nz; *rows_out = rows; *cols_out = cols; *vals_out = vals; *rhs_out = rhs;
	return 0;
}

void lcgb200_free_host(void* p) { std::free(p); }

// host COO triplets (row-sorted) -> device CSR handle: H2D of the triplets, lcgb200_coo2csr on the device, lcgb200_csr_create
int lcgb200_csr_create_from_coo(lcgb200_csr_t* out, int n, int nnz, const int* rows, const int* cols, const void* vals, int value_type, unsigned flags)
{
	if (!out || !rows || !cols || !vals) return LCGB200_INVALID_POINTER;
	if (n <= 0 || nnz < 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	const size_t vs = value_type == LCGB200_REAL ? sizeof(double) : 2 * sizeof(double);
	int *d_rows = nullptr, *d_cols = nullptr, *d_rp = nullptr; void* d_vals = nullptr;
	int rc = LCGB200_UNKNOWN_ERROR;
	if (cudaMalloc((void**)&d_rows, sizeof(int) * (size_t)std::max(nnz, 1)) == cudaSuccess && cudaMalloc((void**)&d_cols, sizeof(int) * (size_t)std::max(nnz, 1)) == cudaSuccess &&
	    cudaMalloc((void**)&d_rp, sizeof(int) * ((size_t)n + 1)) == cudaSuccess && cudaMalloc(&d_vals, vs * (size_t)std::max(nnz, 1)) == cudaSuccess)
	{
		cudaMemcpy(d_rows, rows, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice);
		cudaMemcpy(d_cols, cols, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice);
		cudaMemcpy(d_vals, vals, vs * (size_t)nnz, cudaMemcpyHostToDevice);
		rc = lcgb200_coo2csr(d_rows, nnz, n, d_rp, nullptr);
		if (rc == 0) { cudaDeviceSynchronize(); rc = lcgb200_csr_create(out, n, nnz, d_rp, d_cols, d_vals, value_type, LCGB200_DEVICE, flags); }
	}
	cudaFree(d_rows); cudaFree(d_cols); cudaFree(d_rp); cudaFree(d_vals);
	return rc;
}

}  // extern "C"
