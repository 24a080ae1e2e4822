// csr.cuh — built-in CSR operator: device layout, row tiles, and the streaming SpMV kernel with a fused
// per-row epilogue (dot products / vector updates) and the same single-pass grid reduction as k_vec.
//
// Layout in HBM: int32 row_ptr[n+1], int32 col[nnz (+pad)], T val[nnz (+pad)], base 0, rows in
// order — exactly the arrays the reference's samples hand to cusparseCreateCsr (sample8.cu:172-173) — plus one
// int4 {row_begin, row_end, nnz_begin_aligned, nnz_end} per row tile.  A tile is a run of consecutive rows
// whose non-zeros (from the 4-aligned start) fit the shared-memory staging buffer.
//
// Kernel: a persistent grid (multiple of 148 SMs) walks the tiles round-robin, so that at any moment all CTAs
// work inside one narrow band of rows and the gathered x / written y of that band stay L2-resident.  Per tile:
//   phase 1  col/val of the tile are streamed HBM -> shared memory with 128-bit coalesced loads that do not
//            allocate in L1 (the matrix is read exactly once);
//   phase 2  LPR lanes per row (1..32, picked from the mean row length) walk their row out of shared memory
//            and gather x[col] through L1/L2; lanes of a warp sit on consecutive rows, so for banded/stencil
//            matrices a warp-wide gather touches 2-3 lines instead of ~10;
//   epilogue the row result feeds the fused reductions (p.Ap, r0~.Ap, As.s, As.As, ...) without re-reading y.
#pragma once
#include "common.cuh"

namespace lcgb200 {

constexpr int kTileNnzReal = 2048;   // staged non-zeros per tile: 16 KB val + 8 KB col
constexpr int kTileNnzCplx = 2048;   // 32 KB val + 8 KB col
constexpr int kTileRows = 1024;      // max rows per tile (bounds the row_ptr slice in shared memory)

template <class T>
struct CsrDev {
	int n_rows = 0, n_cols = 0, nnz = 0, n_tiles = 0, lpr = 1;
	const int* row_ptr = nullptr;
	const int* col = nullptr;
	const T* val = nullptr;
	const int4* tiles = nullptr;
};

template <class T> struct TileCfg;
template <> struct TileCfg<double> { static constexpr int NNZ = kTileNnzReal; };
template <> struct TileCfg<double2> { static constexpr int NNZ = kTileNnzCplx; };

__device__ __forceinline__ double mulacc(double acc, double a, double x) { return fma(a, x, acc); }
__device__ __forceinline__ double2 mulacc(double2 acc, double2 a, double2 x)
{
	acc.x = fma(a.x, x.x, acc.x); acc.x = fma(-a.y, x.y, acc.x);
	acc.y = fma(a.x, x.y, acc.y); acc.y = fma(a.y, x.x, acc.y);
	return acc;
}
__device__ __forceinline__ double tzero(double) { return 0.0; }
__device__ __forceinline__ double2 tzero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double tconj(double a) { return a; }
__device__ __forceinline__ double2 tconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double tadd(double a, double b) { return a + b; }
__device__ __forceinline__ double2 tadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double tshfl_xor(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double2 tshfl_xor(double2 v, int o)
{
	return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ double tldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ double2 tldg(const double2* p) { return __ldg(p); }

// stage 4 consecutive non-zeros (k is a multiple of 4) into shared memory
__device__ __forceinline__ void stage4(const int* col, const double* val, int k, int* scol, double* sval, int s)
{
	int4 c = ldg_stream_i4(reinterpret_cast<const int4*>(col + k));
	double2 v0 = ldg_stream_d2(reinterpret_cast<const double2*>(val + k));
	double2 v1 = ldg_stream_d2(reinterpret_cast<const double2*>(val + k + 2));
	*reinterpret_cast<int4*>(scol + s) = c;
	*reinterpret_cast<double2*>(sval + s) = v0;
	*reinterpret_cast<double2*>(sval + s + 2) = v1;
}
__device__ __forceinline__ void stage4(const int* col, const double2* val, int k, int* scol, double2* sval, int s)
{
	int4 c = ldg_stream_i4(reinterpret_cast<const int4*>(col + k));
	double2 v0 = ldg_stream_d2(val + k), v1 = ldg_stream_d2(val + k + 1);
	double2 v2 = ldg_stream_d2(val + k + 2), v3 = ldg_stream_d2(val + k + 3);
	*reinterpret_cast<int4*>(scol + s) = c;
	sval[s] = v0; sval[s + 1] = v1; sval[s + 2] = v2; sval[s + 3] = v3;
}

// Epi interface:
//   static constexpr int NRED;
//   __device__ void begin(const DevState*);
//   __device__ void row(int i, T yi, T xi, double* acc);   called once per row by one lane (may write vectors)
//   __device__ void finish(DevState*, const double* tot);
template <class T, int LPR, bool CONJ, class Epi>
__global__ void __launch_bounds__(kThreads) k_spmv(CsrDev<T> A, const T* __restrict__ x, T* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	if (st_done(st)) return;
	constexpr int TN = TileCfg<T>::NNZ;
	__shared__ __align__(16) T sval[TN + 8];
	__shared__ __align__(16) int scol[TN + 8];
	__shared__ int srow[kTileRows + 1];
	__shared__ T s_long[kThreads / 32];

	Epi epi = epi_in;
	epi.begin(st);
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;

	constexpr int NG = kThreads / LPR;          // row groups per block
	const int group = threadIdx.x / LPR, lane = threadIdx.x % LPR;

	for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x)
	{
		const int4 td = A.tiles[tile];
		const int r0 = td.x, nrows = td.y - td.x, k0 = td.z, k1 = td.w;
		if (k1 - k0 > TN)
		{	// one row longer than the staging buffer: stream it straight from global memory
			T part = tzero(T());
			for (int k = A.row_ptr[r0] + threadIdx.x; k < k1; k += kThreads)
			{
				T a = A.val[k]; if (CONJ) a = tconj(a);
				part = mulacc(part, a, tldg(x + A.col[k]));
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) part = tadd(part, tshfl_xor(part, o));
			if ((threadIdx.x & 31) == 0) s_long[threadIdx.x >> 5] = part;
			__syncthreads();
			if (threadIdx.x == 0)
			{
				T tot = s_long[0];
				for (int w = 1; w < kThreads / 32; w++) tot = tadd(tot, s_long[w]);
				y[r0] = tot;
				epi.row(r0, tot, tldg(x + r0), acc);
			}
			__syncthreads();
			continue;
		}
		// phase 1: stage the tile
		for (int i = threadIdx.x; i <= nrows; i += kThreads) srow[i] = A.row_ptr[r0 + i] - k0;
		for (int k = k0 + 4 * threadIdx.x; k < k1; k += 4 * kThreads) stage4(A.col, A.val, k, scol, sval, k - k0);
		__syncthreads();
		// phase 2: rows
		for (int rb = 0; rb < nrows; rb += NG)
		{
			const int r = rb + group;
			int kb = 0, ke = 0;
			if (r < nrows) { kb = srow[r]; ke = srow[r + 1]; }
			T sum = tzero(T());
#pragma unroll 4
			for (int j = kb + lane; j < ke; j += LPR)
			{
				T a = sval[j]; if (CONJ) a = tconj(a);
				sum = mulacc(sum, a, tldg(x + scol[j]));
			}
#pragma unroll
			for (int o = LPR / 2; o > 0; o >>= 1) sum = tadd(sum, tshfl_xor(sum, o));
			if (lane == 0 && r < nrows)
			{
				y[r0 + r] = sum;
				epi.row(r0 + r, sum, tldg(x + r0 + r), acc);
			}
		}
		__syncthreads();
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { for (int r = 0; r < Epi::NRED; r++) st->red[r] = tot[r]; }
			else epi.finish(st, tot);
		}
	}
}

// Adaptor: run an SpMV epilogue as a plain vector kernel over an already computed y (user-callback operators).
template <class T, class Epi>
struct RowEpilogueOp {
	static constexpr int NRED = Epi::NRED;
	static constexpr int W = 1;
	Epi epi; const T* x; const T* y;
	__device__ bool active(const DevState*) const { return true; }
	__device__ void begin(const DevState* st) { epi.begin(st); }
	template <int V> __device__ void elem(size_t i, double* acc) { epi.row((int)i, y[i], x[i], acc); }
	__device__ void finish(DevState* st, const double* tot) { epi.finish(st, tot); }
};

// epilogue that does nothing (plain y = A x)
template <class T>
struct EpiNone {
	static constexpr int NRED = 0;
	static constexpr bool ACTIVE = false;   // nothing to run after a user-callback SpMV
	__device__ void begin(const DevState*) {}
	__device__ void row(int, T, T, double*) {}
	__device__ void finish(DevState*, const double*) {}
};

inline int spmv_grid(int n_tiles)
{
	int g = n_tiles < kMaxBlocks ? n_tiles : kMaxBlocks;
	return g < 1 ? 1 : g;
}

// launch with the lanes-per-row variant recorded in the handle
template <class T, bool CONJ, class Epi>
inline void launch_spmv(const CsrDev<T>& A, const T* x, T* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	const int grid = spmv_grid(A.n_tiles);
	switch (A.lpr)
	{
		case 1: k_spmv<T, 1, CONJ, Epi><<<grid, kThreads, 0, s>>>(A, x, y, epi, st, partials); break;
		case 2: k_spmv<T, 2, CONJ, Epi><<<grid, kThreads, 0, s>>>(A, x, y, epi, st, partials); break;
		case 4: k_spmv<T, 4, CONJ, Epi><<<grid, kThreads, 0, s>>>(A, x, y, epi, st, partials); break;
		case 8: k_spmv<T, 8, CONJ, Epi><<<grid, kThreads, 0, s>>>(A, x, y, epi, st, partials); break;
		case 16: k_spmv<T, 16, CONJ, Epi><<<grid, kThreads, 0, s>>>(A, x, y, epi, st, partials); break;
		default: k_spmv<T, 32, CONJ, Epi><<<grid, kThreads, 0, s>>>(A, x, y, epi, st, partials); break;
	}
}

}  // namespace lcgb200
