// fused_small.cuh — cache-resident systems (the reference's own sample cases, data/case_10K_*: n = 10^4, ~1 MB): there
// the iteration is bound by launch and grid-reduction latency, not by HBM.  One cooperative kernel runs several whole
// iterations — the solver's list of phases (SpMV+dot, fused updates, direction) — with grid-wide barriers in
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

// A phase that ends in a reduction needs TWO things from the grid: the totals (every block's partial) and a barrier (the
// next phase reads what other blocks wrote and the scalars the epilogue produced).  Both are one round: a block's arrival
// at the reduction ticket IS its arrival at the barrier; the block that arrives last totals the partials, runs the scalar
// epilogue and opens the gate (DevState::gate, a monotonic counter) the others spin on — instead of a reduction round
// followed by a separate grid.sync() round (each ~2 us on 148-296 blocks, which is what an iteration of a 10^4-row system is
// made of).  Returns after the gate has opened; `gate` is the block's private copy of the counter.
__device__ __forceinline__ void gate_open(DevState* st, int next) { __threadfence(); st_release_gpu_s32(&st->gate, next); }
__device__ __forceinline__ void gate_wait(DevState* st, int& gate)
{
	if (threadIdx.x == 0) { const int g0 = gate; while (ld_acquire_gpu_s32(&st->gate) == g0) { } }
	gate++;
	__syncthreads();
}

template <class Op>
__device__ __forceinline__ void phase_vec(Op& op, size_t n, DevState* st, double* partials, int& gate)
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
		{
			if ((threadIdx.x & 31) == 0) { op.finish(st, tot); gate_open(st, gate + 1); }
		}
		gate_wait(st, gate);
	}
}

// y = op(A) x straight out of L2: `lpr` lanes per row (run-time here), the same per-lane accumulation order and
// butterfly as k_spmv, so a row sum is bitwise what the streaming kernel produces
template <class T, bool CONJ, class Epi>
__device__ __forceinline__ void phase_spmv(const CsrDev<T>& A, const T* x, T* y, Epi& epi, DevState* st, double* partials, int& gate)
{
	epi.begin(st);
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;
	const int lpr = A.lpr;
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const int ngroups = (gridDim.x * blockDim.x) / lpr;
	const int group = gtid / lpr, lane = gtid % lpr;
	for (int base = 0; base < A.n_rows; base += ngroups)
	{
		const int row = base + group;
		const bool valid = row < A.n_rows;
		int kb = 0, ke = 0;
		if (valid) { kb = A.row_ptr[row]; ke = A.row_ptr[row + 1]; }
		T sum = tzero(T());
		for (int j0 = kb + lane; j0 < ke; j0 += lpr * kGatherUnroll)
		{
			int cidx[kGatherUnroll]; T a[kGatherUnroll], xv[kGatherUnroll];
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++)
			{
				const int j = j0 + u * lpr;
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
		for (int o = lpr / 2; o > 0; o >>= 1) sum = tadd(sum, tshfl_xor(sum, o));
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
		{
			if ((threadIdx.x & 31) == 0) { epi.finish(st, tot); gate_open(st, gate + 1); }
		}
		gate_wait(st, gate);
	}
}

// phase descriptors: what one launch of the streaming path would have been
template <class T, bool CONJ, class Epi>
struct SpmvPhase {
	CsrDev<T> A; const T* x; T* y; Epi epi;
	static constexpr bool GATED = Epi::NRED > 0;   // the phase ends in a reduction whose gate is its barrier
	__device__ void run(DevState* st, double* partials, int& gate) { Epi e = epi; phase_spmv<T, CONJ, Epi>(A, x, y, e, st, partials, gate); }
	int rows() const { return A.n_rows * A.lpr; }
};
template <class Op>
struct VecPhase {
	Op op; size_t n;
	static constexpr bool GATED = Op::NRED > 0;
	__device__ void run(DevState* st, double* partials, int& gate) { Op o = op; phase_vec(o, n, st, partials, gate); }
	int rows() const { return (int)(n / Op::W); }
};

template <class P>
__device__ __forceinline__ void run_phase(P& p, DevState* st, double* partials, cg::grid_group& grid, int& gate)
{
	// like the streaming kernels, nothing happens once the solve is over (`done` only changes right before a barrier every
	// block has passed, so the decision is uniform across the grid)
	if (st_done(st)) return;
	p.run(st, partials, gate);
	if (!P::GATED) grid.sync();
}

// `iters` iterations of the phase list per launch (cooperative: the whole grid is co-resident)
template <class... Ph>
__global__ void __launch_bounds__(kThreads) k_fused(DevState* st, double* partials, int iters, Ph... ph)
{
	cg::grid_group grid = cg::this_grid();
	int gate = ld_acquire_gpu_s32(&st->gate);   // no block can open a gate before every block has arrived at the first reduction
	for (int it = 0; it < iters; it++)
	{
		if (st_done(st)) return;   // `done` only changes right before a barrier every block has passed: a uniform decision
		(run_phase(ph, st, partials, grid, gate), ...);
	}
}

int coop_grid_limit(const void* kernel, int block);   // co-resident blocks of a kernel on the current device (engine.cu)

template <class... Ph>
inline cudaError_t launch_fused(DevState* st, double* partials, int iters, cudaStream_t s, Ph... ph)
{
	auto kern = k_fused<Ph...>;
	static int limits[64] = {0};   // per instantiation and device
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
	if (!limits[dev]) limits[dev] = coop_grid_limit((const void*)kern, kThreads);
	const int limit = limits[dev];
	int work = 1;
	int each[] = {ph.rows()...};
	for (int w : each) work = w > work ? w : work;
	int grid = (work + kThreads - 1) / kThreads;
	grid = grid < 1 ? 1 : (grid > limit ? limit : grid);
	void* args[] = {&st, &partials, &iters, &ph...};
	return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), args, 0, s);
}

}  // namespace lcgb200
