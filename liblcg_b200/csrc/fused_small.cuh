// fused_small.cuh — cache-resident systems (the reference's own sample cases, data/case_10K_*: n = 10^4, ~1 MB): there
// the iteration is bound by launch and grid-reduction latency, not by HBM.  One cooperative kernel runs several whole
// iterations of the 3-step shape  SpMV(+dot) -> update(+norms, loop head) -> direction  with grid-wide barriers in
// between, re-using the very same functors (epilogue / Op::elem / Op::finish) the streaming kernels use, so the
// arithmetic, the scalar epilogues and the reference's loop-head control are identical by construction.
//
// Memory-model notes: vectors written in one phase are read in the next by other blocks, so this kernel uses plain
// (coherent) loads only — no ld.global.nc — and relies on grid.sync() for ordering.  DevState scalars are re-read
// (Op::begin) after each barrier.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "csr.cuh"

namespace lcgb200 {

namespace cg = cooperative_groups;

constexpr int kSmallRows = 65536;      // systems up to this many rows ...
constexpr int kSmallNnz = 1 << 21;     // ... and non-zeros take the fused path (matrix + vectors stay L2-resident)

template <class Op>
__device__ __forceinline__ void phase_vec(Op& op, size_t n, DevState* st, double* partials)
{
	if (!op.active(st)) return;   // uniform across the grid: st only changes right before a barrier everyone passed
	op.begin(st);
	double acc[Op::NRED > 0 ? Op::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Op::NRED > 0 ? Op::NRED : 1); r++) acc[r] = 0.0;
	constexpr int W = Op::W;
	const size_t npack = n / W;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npack; p += stride) op.template elem<W>(p * W, acc);
	if (W > 1)
	{
		const size_t tail = npack * W + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
		if (tail < n) op.template elem<1>(tail, acc);
	}
	if (Op::NRED > 0)
	{
		double tot[Op::NRED > 0 ? Op::NRED : 1];
		if (grid_reduce<(Op::NRED > 0 ? Op::NRED : 1)>(acc, partials, &st->ticket, tot))
			if ((threadIdx.x & 31) == 0) op.finish(st, tot);
	}
}

// y = A x straight out of L2: LPR lanes per row, the same per-lane accumulation order and butterfly as k_spmv
template <class T, int LPR, bool CONJ, class Epi>
__device__ __forceinline__ void phase_spmv(const CsrDev<T>& A, const T* x, T* y, Epi& epi, DevState* st, double* partials)
{
	epi.begin(st);
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const int ngroups = (gridDim.x * blockDim.x) / LPR;
	const int group = gtid / LPR, lane = gtid % LPR;
	for (int base = 0; base < A.n_rows; base += ngroups)
	{
		const int row = base + group;
		const bool valid = row < A.n_rows;
		int kb = 0, ke = 0;
		if (valid) { kb = A.row_ptr[row]; ke = A.row_ptr[row + 1]; }
		T sum = tzero(T());
		for (int j0 = kb + lane; j0 < ke; j0 += LPR * kGatherUnroll)
		{
			int cidx[kGatherUnroll]; T a[kGatherUnroll], xv[kGatherUnroll];
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++)
			{
				const int j = j0 + u * LPR;
				const bool ok = j < ke;
				cidx[u] = ok ? A.col[j] : -1;
				a[u] = ok ? A.val[j] : tzero(T());
			}
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++) xv[u] = cidx[u] >= 0 ? x[cidx[u]] : tzero(T());
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++)
			{
				if (CONJ) a[u] = tconj(a[u]);
				sum = mulacc(sum, a[u], xv[u]);
			}
		}
#pragma unroll
		for (int o = LPR / 2; o > 0; o >>= 1) sum = tadd(sum, tshfl_xor(sum, o));
		if (lane == 0 && valid)
		{
			y[row] = sum;
			epi.row(row, sum, x[row], acc);
		}
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
			if ((threadIdx.x & 31) == 0) epi.finish(st, tot);
	}
}

// `iters` iterations of  SpMV(x -> y, Epi) ; Op1 ; Op2  per launch (cooperative: the whole grid is co-resident)
template <class T, int LPR, class Epi, class Op1, class Op2>
__global__ void __launch_bounds__(kThreads) k_fused3(CsrDev<T> A, const T* x, T* y, Epi epi_in, Op1 op1_in, Op2 op2_in, size_t n,
	DevState* st, double* partials, int iters)
{
	cg::grid_group grid = cg::this_grid();
	for (int it = 0; it < iters; it++)
	{
		if (st_done(st)) return;   // set only right before a barrier every block has passed: a uniform decision
		{ Epi epi = epi_in; phase_spmv<T, LPR, false, Epi>(A, x, y, epi, st, partials); }
		grid.sync();
		if (!st_done(st)) { Op1 op1 = op1_in; phase_vec(op1, n, st, partials); }
		grid.sync();
		if (!st_done(st)) { Op2 op2 = op2_in; phase_vec(op2, n, st, partials); }
		grid.sync();
	}
}

int coop_grid_limit(const void* kernel, int block);   // co-resident blocks of a kernel on the current device (engine.cu)

template <class T, int LPR, class Epi, class Op1, class Op2>
inline cudaError_t launch_fused3_lpr(const CsrDev<T>& A, const T* x, T* y, const Epi& epi, const Op1& op1, const Op2& op2, size_t n,
	DevState* st, double* partials, int iters, cudaStream_t s)
{
	auto kern = k_fused3<T, LPR, Epi, Op1, Op2>;
	static int limit = 0;   // per instantiation
	if (!limit) limit = coop_grid_limit((const void*)kern, kThreads);
	long long want = ((long long)A.n_rows * LPR + kThreads - 1) / kThreads;
	int grid = (int)(want < 1 ? 1 : (want > limit ? limit : want));
	CsrDev<T> Ac = A; const T* xc = x; T* yc = y; Epi e = epi; Op1 o1 = op1; Op2 o2 = op2; size_t nc = n; DevState* stc = st; double* pc = partials; int ic = iters;
	void* args[] = {&Ac, &xc, &yc, &e, &o1, &o2, &nc, &stc, &pc, &ic};
	return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), args, 0, s);
}

template <class T, class Epi, class Op1, class Op2>
inline cudaError_t launch_fused3(const CsrDev<T>& A, const T* x, T* y, const Epi& epi, const Op1& op1, const Op2& op2, size_t n,
	DevState* st, double* partials, int iters, cudaStream_t s)
{
	switch (A.lpr)
	{
		case 1: return launch_fused3_lpr<T, 1>(A, x, y, epi, op1, op2, n, st, partials, iters, s);
		case 2: return launch_fused3_lpr<T, 2>(A, x, y, epi, op1, op2, n, st, partials, iters, s);
		case 4: return launch_fused3_lpr<T, 4>(A, x, y, epi, op1, op2, n, st, partials, iters, s);
		case 8: return launch_fused3_lpr<T, 8>(A, x, y, epi, op1, op2, n, st, partials, iters, s);
		case 16: return launch_fused3_lpr<T, 16>(A, x, y, epi, op1, op2, n, st, partials, iters, s);
		default: return launch_fused3_lpr<T, 32>(A, x, y, epi, op1, op2, n, st, partials, iters, s);
	}
}

}  // namespace lcgb200
