// exact.cuh — kernels of the reference-order build (VariantExact; common.cuh, "arithmetic variants").
//
// What has to match for a GPU solve to be bit-identical to the reference's CPU solve:
//   * element-wise updates: same expression, every product and every sum rounded on its own (the translation unit is
//     compiled with -fmad=false and fma() is a multiply followed by an add) — the reference's loops are
//     `#pragma omp parallel for` over independent elements (lcg.cpp:237-242 ...), so their order does not matter;
//   * row sums of the SpMV: one thread per row adds the row's products left to right (the caller's Ax callback in the
//     reference's tests and samples is a plain CSR loop);
//   * dot products: lcg_dot / clcg_inner / clcg_dot are SERIAL loops `ret += a[i]*b[i]` (algebra.cpp:154-163,
//     lcg_complex.cpp:143-167).  Every element's term is computed in parallel and stored (terms[slot][i]); one thread per
//     reduction slot then adds the terms left to right in a second, single-block kernel which also runs the step's scalar
//     epilogue (the loop head included).  O(n) dependent additions per reduction: this build is a verification mode for the
//     cache-resident sample systems (10^3-10^4 rows, tens of microseconds per reduction), not a fast path.
// Plain launches only (no programmatic dependent launch, no cooperative kernels); single GPU, double precision.
#pragma once
#include "common.cuh"
#include "csr.cuh"

namespace lcgb200 {

// element-wise step: Op::elem<1> per element with a zeroed accumulator, whose content afterwards is the element's term
template <class Op>
__global__ void __launch_bounds__(kThreads) kx_vec(Op op_in, size_t n, DevState* st, double* terms, size_t stride)
{
	if (st_done(st)) return;
	Op op = op_in;
	if (!op.active(st)) return;
	op.begin(st);
	constexpr int NR = Op::NRED > 0 ? Op::NRED : 1;
	const size_t step = (size_t)gridDim.x * kThreads;
	for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += step)
	{
		double acc[NR];
#pragma unroll
		for (int r = 0; r < NR; r++) acc[r] = 0.0;
		op.template elem<1>(i, acc);
		if (Op::NRED > 0)
		{
#pragma unroll
			for (int r = 0; r < NR; r++) terms[(size_t)r * stride + i] = acc[r];
		}
	}
}

// y = op(A) x, one thread per row, products added left to right; the fused per-row epilogue leaves its terms like kx_vec
template <class T, bool CONJ, class Epi>
__global__ void __launch_bounds__(kThreads) kx_spmv(CsrDev<T> A, const T* x, T* y, Epi epi_in, DevState* st, double* terms, size_t stride)
{
	if (st_done(st)) return;
	Epi epi = epi_in;
	epi.begin(st);
	constexpr int NR = Epi::NRED > 0 ? Epi::NRED : 1;
	const int step = (int)(gridDim.x * kThreads);
	for (int row = (int)(blockIdx.x * kThreads + threadIdx.x); row < A.n_rows; row += step)
	{
		T sum = tzero(T());
		const int ke = A.row_ptr[row + 1];
		for (int k = A.row_ptr[row]; k < ke; k++)
		{
			T a = A.val[k];
			if (CONJ) a = tconj(a);
			sum = mulacc(sum, a, x[A.col[k]]);
		}
		y[row] = sum;
		double acc[NR];
#pragma unroll
		for (int r = 0; r < NR; r++) acc[r] = 0.0;
		epi.row(row, sum, x[row], acc);
		if (Epi::NRED > 0)
		{
#pragma unroll
			for (int r = 0; r < NR; r++) terms[(size_t)r * stride + row] = acc[r];
		}
	}
}

// serial totals + scalar epilogue: thread r adds slot r's terms in index order, starting from 0.0 like the reference's loops
template <class Fin, int NRED>
__global__ void __launch_bounds__(32) kx_total(Fin f_in, size_t n, DevState* st, const double* terms, size_t stride)
{
	if (st_done(st)) return;
	Fin f = f_in;
	if (!f.active(st)) return;
	f.begin(st);
	__shared__ double tot[kMaxRed];
	if ((int)threadIdx.x < NRED)
	{
		const double* t = terms + (size_t)threadIdx.x * stride;
		double s = 0.0;
		size_t i = 0;
		for (; i + 8 <= n; i += 8)
		{	// the loads of eight terms are issued together, the additions stay in order
			double v[8];
#pragma unroll
			for (int u = 0; u < 8; u++) v[u] = __ldcg(t + i + u);
#pragma unroll
			for (int u = 0; u < 8; u++) s = __dadd_rn(s, v[u]);
		}
		for (; i < n; i++) s = __dadd_rn(s, __ldcg(t + i));
		tot[threadIdx.x] = s;
	}
	__syncthreads();
	if (threadIdx.x == 0) f.finish(st, tot);
}

// an SpMV epilogue seen as the "finisher" of kx_total
template <class Epi>
struct EpiFinisher {
	Epi epi;
	__device__ bool active(const DevState*) const { return true; }
	__device__ void begin(const DevState* st) { epi.begin(st); }
	__device__ void finish(DevState* st, const double* tot) { epi.finish(st, tot); }
};

template <class Op>
inline void launch_exact_vec(const Op& op, size_t n, DevState* st, double* terms, size_t stride, cudaStream_t s)
{
	size_t blocks = (n + kThreads - 1) / kThreads;
	if (blocks < 1) blocks = 1;
	if (blocks > (size_t)kMaxBlocks) blocks = kMaxBlocks;
	kx_vec<Op><<<(unsigned)blocks, kThreads, 0, s>>>(op, n, st, terms, stride);
	if (Op::NRED > 0) kx_total<Op, (Op::NRED > 0 ? Op::NRED : 1)><<<1, 32, 0, s>>>(op, n, st, terms, stride);
}

template <class T, bool CONJ, class Epi>
inline void launch_exact_spmv(const CsrDev<T>& A, const T* x, T* y, const Epi& epi, DevState* st, double* terms, size_t stride, cudaStream_t s)
{
	size_t blocks = ((size_t)A.n_rows + kThreads - 1) / kThreads;
	if (blocks < 1) blocks = 1;
	if (blocks > (size_t)kMaxBlocks) blocks = kMaxBlocks;
	kx_spmv<T, CONJ, Epi><<<(unsigned)blocks, kThreads, 0, s>>>(A, x, y, epi, st, terms, stride);
	if (Epi::NRED > 0)
		kx_total<EpiFinisher<Epi>, (Epi::NRED > 0 ? Epi::NRED : 1)><<<1, 32, 0, s>>>(EpiFinisher<Epi>{epi}, (size_t)A.n_rows, st, terms, stride);
}

}  // namespace lcgb200
