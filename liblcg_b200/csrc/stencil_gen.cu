// stencil_gen.cu — device-side generator of the synthetic systems of SURVEY.md §8(d) (bench / large tests).
// Produces, for global rows [row0, row1) of a g^3 grid, exactly the CSR arrays liblcg_b200/stencil.py defines
// (ascending columns, Dirichlet truncation) and b = A x* with x*[i] = uint32(i*2654435761)/2^32 summed left to
// right.  Column indices are written as (global column - col_offset) so that a single-GPU caller passes 0.
#include "common.cuh"
#include "../../include/lcgb200.h"

namespace lcgb200 {

__device__ __forceinline__ double xstar(long long i)
{
	unsigned int h = (unsigned int)((unsigned long long)i * 2654435761ull);
	return (double)h / 4294967296.0;
}

// visit the stencil entries of `row` in ascending column order; f(col, val)
template <class F>
__device__ __forceinline__ int visit_row(int kind, int g, long long row, F f)
{
	const long long gg = (long long)g * g;
	const int x = (int)(row % g), y = (int)((row / g) % g), z = (int)(row / gg);
	int cnt = 0;
	if (kind == 1)
	{
		for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++)
		{
			const int zz = z + dz, yy = y + dy, xx = x + dx;
			if (zz < 0 || zz >= g || yy < 0 || yy >= g || xx < 0 || xx >= g) continue;
			f(((long long)zz * g + yy) * g + xx, (dz == 0 && dy == 0 && dx == 0) ? 26.0 : -1.0, cnt);
			cnt++;
		}
		return cnt;
	}
	const double gx = kind == 2 ? 0.5 : 0.0, gy = kind == 2 ? 0.25 : 0.0, gz = kind == 2 ? 0.125 : 0.0;
	if (z > 0) { f(row - gg, -1.0 - gz, cnt); cnt++; }
	if (y > 0) { f(row - g, -1.0 - gy, cnt); cnt++; }
	if (x > 0) { f(row - 1, -1.0 - gx, cnt); cnt++; }
	f(row, 6.0, cnt); cnt++;
	if (x < g - 1) { f(row + 1, -1.0 + gx, cnt); cnt++; }
	if (y < g - 1) { f(row + g, -1.0 + gy, cnt); cnt++; }
	if (z < g - 1) { f(row + gg, -1.0 + gz, cnt); cnt++; }
	return cnt;
}

// closed form of the number of entries in rows [0, row) — lets every thread place its row without a scan
__device__ __host__ inline long long nnz_before(int kind, int g, long long row)
{
	const long long gg = (long long)g * g;
	const long long z = row / gg, rem = row % gg, y = rem / g, x = rem % g;
	if (kind == 1)
	{
		// valid offsets along one axis at coordinate c: 3 inside, 2 at either end (1 when g == 1)
		auto w = [&](long long c) { return g == 1 ? 1LL : ((c == 0 || c == g - 1) ? 2LL : 3LL); };
		auto S = [&](long long k) { return g == 1 ? k : (3 * k - (k >= 1 ? 1 : 0) - (k == g ? 1 : 0)); };	// sum_{c<k} w(c)
		const long long full = S(g);
		return full * full * S(z) + full * S(y) * w(z) + S(x) * w(y) * w(z);
	}
	// 7-point: a row has 1 + c(x) + c(y) + c(z) entries, c(.) = number of valid neighbours along that axis
	auto c = [&](long long q) { return (long long)(q > 0) + (long long)(q < g - 1); };
	auto C = [&](long long k) { return g == 1 ? 0LL : (2 * k - (k >= 1 ? 1 : 0) - (k == g ? 1 : 0)); };	// sum_{q<k} c(q)
	long long total = z * (gg + 4LL * g * (g - 1)) + gg * C(z);
	total += y * ((long long)g * (1 + c(z)) + 2LL * (g - 1)) + (long long)g * C(y);
	total += x * (1 + c(z) + c(y)) + C(x);
	return total;
}

__global__ void k_gen_stencil(int kind, int g, long long row0, long long row1, long long base, int* rp, int* ci, double* v, long long col_offset)
{
	const long long row = row0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (row > row1) return;
	const long long start = nnz_before(kind, g, row) - base;
	rp[row - row0] = (int)start;
	if (row == row1) return;
	visit_row(kind, g, row, [&](long long c, double val, int j) { ci[start + j] = (int)(c - col_offset); v[start + j] = val; });
}

__global__ void k_gen_rhs(int kind, int g, long long row0, long long row1, double* b)
{
	const long long row = row0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= row1) return;
	double acc = 0.0;
	visit_row(kind, g, row, [&](long long c, double val, int) { acc = __dadd_rn(acc, __dmul_rn(val, xstar(c))); });
	b[row - row0] = acc;
}

}  // namespace lcgb200

using namespace lcgb200;

extern "C" int lcgb200_gen_stencil(int kind, int g, long long row0, long long row1, int* rp, int* ci, double* v,
	long long col_offset, long long* nnz_out, void* stream)
{
	if (kind < 0 || kind > 2 || g <= 0 || row0 < 0 || row1 < row0 || row1 > (long long)g * g * g) return LCGB200_INVILAD_VARIABLE_SIZE;
	const long long base = nnz_before(kind, g, row0);
	const long long nnz = nnz_before(kind, g, row1) - base;
	if (nnz_out) *nnz_out = nnz;
	if (!rp || !ci || !v) return 0;
	if (nnz >= 2147483647LL) return LCGB200_INVILAD_VARIABLE_SIZE;
	const long long rows = row1 - row0 + 1;
	k_gen_stencil<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, g, row0, row1, base, rp, ci, v, col_offset);
	return cudaGetLastError() == cudaSuccess ? 0 : LCGB200_UNKNOWN_ERROR;
}

extern "C" int lcgb200_gen_rhs(int kind, int g, long long row0, long long row1, double* b, void* stream)
{
	if (kind < 0 || kind > 2 || g <= 0 || row0 < 0 || row1 < row0 || row1 > (long long)g * g * g || !b) return LCGB200_INVILAD_VARIABLE_SIZE;
	const long long rows = row1 - row0;
	if (rows == 0) return 0;
	k_gen_rhs<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, g, row0, row1, b);
	return cudaGetLastError() == cudaSuccess ? 0 : LCGB200_UNKNOWN_ERROR;
}
