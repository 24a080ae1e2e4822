// datastep.cu — the step in front of the path: liblcg's binary COO fixtures -> CSR on the
// device, without cuSPARSE.  The reference's samples read the file on the host (sample8.cu:30-64), copy the triplets
// over and call cusparseXcoo2csr on the row-sorted indices (sample8.cu:169); lcgb200_coo2csr is that call.
#include "common.cuh"
#include "../../include/lcgb200.h"
#include <cuComplex.h>
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

// ---- the element-wise device helpers of algebra_cuda.h / lcg_complex_cuda.h (what the reference's samples build their
// Jacobi Mx callbacks from: sample10.cu:117,193).  One grid-stride kernel per operation; the complex products and
// quotients are cuCmul / cuCdiv(f), the functions the reference's kernels call (lcg_complex_cuda.cu:62-101).
struct HelpMul { __device__ double operator()(double a, double b) const { return a * b; }
	__device__ cuDoubleComplex operator()(cuDoubleComplex a, cuDoubleComplex b) const { return cuCmul(a, b); }
	__device__ cuComplex operator()(cuComplex a, cuComplex b) const { return cuCmulf(a, b); } };
struct HelpDiv { __device__ double operator()(double a, double b) const { return a / b; }
	__device__ cuDoubleComplex operator()(cuDoubleComplex a, cuDoubleComplex b) const { return cuCdiv(a, b); }
	__device__ cuComplex operator()(cuComplex a, cuComplex b) const { return cuCdivf(a, b); } };
struct HelpConj { __device__ double operator()(double a, double) const { return a; }
	__device__ cuDoubleComplex operator()(cuDoubleComplex a, cuDoubleComplex) const { a.y *= -1.0; return a; }
	__device__ cuComplex operator()(cuComplex a, cuComplex) const { a.y *= -1.0; return a; } };

template <class T, class Op, bool BINARY>
__global__ void k_help_elementwise(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ c, int n)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c[i] = Op()(a[i], BINARY ? b[i] : a[i]);
}

// A_diag[i] = A[i, i]; rows without a diagonal entry keep what A_diag held (algebra_cuda.cu:40-57, lcg_complex_cuda.cu:27-61)
template <class T>
__global__ void k_help_diagonal(const int* __restrict__ rp, const int* __restrict__ ci, const T* __restrict__ v, int n, T* __restrict__ d)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) { d[i] = v[k]; break; }
}

// lcg_set2box_cuda (algebra_cuda.cu:26-38): closed or open bounds make no difference to the result
__global__ void k_help_set2box(const double* __restrict__ low, const double* __restrict__ hig, double* __restrict__ a, int n)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		double v = a[i];
		if (v > hig[i]) v = hig[i];
		if (v < low[i]) v = low[i];
		a[i] = v;
	}
}

template <class T>
int help_elementwise(int op, const void* a, const void* b, void* c, int n, cudaStream_t s)
{
	const int grid = std::max(1, std::min((n + 255) / 256, 148 * 8));
	if (op == 0) k_help_elementwise<T, HelpMul, true><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (T*)c, n);
	else if (op == 1) k_help_elementwise<T, HelpDiv, true><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (T*)c, n);
	else if (op == 2) k_help_elementwise<T, HelpConj, false><<<grid, 256, 0, s>>>((const T*)a, (const T*)a, (T*)c, n);
	else return LCGB200_INVILAD_VARIABLE_SIZE;
	return cudaGetLastError() == cudaSuccess ? 0 : LCGB200_UNKNOWN_ERROR;
}

}  // namespace lcgb200

using namespace lcgb200;

extern "C" {

int lcgb200_vec_elementwise(int op, int value_type, const void* a_dev, const void* b_dev, void* c_dev, int n, void* stream)
{
	if (!a_dev || !c_dev || (op != 2 && !b_dev)) return LCGB200_INVALID_POINTER;
	if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	cudaStream_t s = (cudaStream_t)stream;
	if (value_type == LCGB200_REAL) return help_elementwise<double>(op, a_dev, b_dev, c_dev, n, s);
	if (value_type == LCGB200_COMPLEX) return help_elementwise<cuDoubleComplex>(op, a_dev, b_dev, c_dev, n, s);
	if (value_type == LCGB200_COMPLEX_FLOAT) return help_elementwise<cuComplex>(op, a_dev, b_dev, c_dev, n, s);
	return LCGB200_INVILAD_VARIABLE_SIZE;
}

int lcgb200_diagonal_of_csr(int value_type, const int* row_ptr_dev, const int* col_dev, const void* val_dev, int n, void* diag_dev, void* stream)
{
	if (!row_ptr_dev || !col_dev || !val_dev || !diag_dev) return LCGB200_INVALID_POINTER;
	if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	cudaStream_t s = (cudaStream_t)stream;
	const int grid = std::max(1, std::min((n + 255) / 256, 148 * 8));
	if (value_type == LCGB200_REAL) k_help_diagonal<double><<<grid, 256, 0, s>>>(row_ptr_dev, col_dev, (const double*)val_dev, n, (double*)diag_dev);
	else if (value_type == LCGB200_COMPLEX) k_help_diagonal<cuDoubleComplex><<<grid, 256, 0, s>>>(row_ptr_dev, col_dev, (const cuDoubleComplex*)val_dev, n, (cuDoubleComplex*)diag_dev);
	else if (value_type == LCGB200_COMPLEX_FLOAT) k_help_diagonal<cuComplex><<<grid, 256, 0, s>>>(row_ptr_dev, col_dev, (const cuComplex*)val_dev, n, (cuComplex*)diag_dev);
	else return LCGB200_INVILAD_VARIABLE_SIZE;
	return cudaGetLastError() == cudaSuccess ? 0 : LCGB200_UNKNOWN_ERROR;
}

int lcgb200_set2box(const double* low_dev, const double* hig_dev, double* a_dev, int n, void* stream)
{
	if (!low_dev || !hig_dev || !a_dev) return LCGB200_INVALID_POINTER;
	if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	k_help_set2box<<<std::max(1, std::min((n + 255) / 256, 148 * 8)), 256, 0, (cudaStream_t)stream>>>(low_dev, hig_dev, a_dev, n);
	return cudaGetLastError() == cudaSuccess ? 0 : LCGB200_UNKNOWN_ERROR;
}

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
