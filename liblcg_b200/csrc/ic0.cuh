// ic0.cuh — applying the IC(0) preconditioner on the GPU: z = L^-T L^-1 r as two sparse triangular solves per PCG iteration
// (what the reference's samples do with cusparseSpSV in their Mx callback, sample12.cu:95-105).
//
// Synchronisation-free triangular solve in level order: the rows are listed level by level (ic0_host.h: level_order), a
// persistent grid hands out consecutive positions of that list through one atomic counter, ONE THREAD per row walks its
// off-diagonal entries and, for each, spins until the row it depends on has published its value (a per-row flag holding the
// solve's epoch — never cleared), then x[row] = (b[row] - sum) / diag.  Rows of one level are independent and levels are
// padded to whole warps, so a warp never waits on itself; a row only ever waits on rows at earlier positions, which running
// threads already own: no deadlock, no kernel launch or grid barrier per level.  Latency-bound by construction (the chain of
// levels is serial), like every SpSV.
#pragma once
#include "common.cuh"

namespace lcgb200 {

struct IcCtl { int epoch; unsigned int next; unsigned int ticket; int pad; };

struct IcDev {	// one triangular factor on the device (CSR rows of L: diagonal LAST; rows of U = L^T: diagonal FIRST)
	int n = 0, n_pos = 0, n_levels = 0;
	int* rp = nullptr; int* ci = nullptr; void* val = nullptr; int* order = nullptr; int* ready = nullptr; IcCtl* ctl = nullptr;
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// arithmetic in double / double2 whatever the storage type
__device__ __forceinline__ double ic_ld(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ic_ld(const double2* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ic_ld(const ZF* p) { const float2 t = __ldcg(reinterpret_cast<const float2*>(p)); return make_double2((double)t.x, (double)t.y); }
__device__ __forceinline__ double ic_fma(double a, double x, double s) { return fma(a, x, s); }
__device__ __forceinline__ double2 ic_fma(double2 a, double2 x, double2 s) { return zadd(s, zmul(a, x)); }
__device__ __forceinline__ double ic_fin(double b, double s, double d) { return (b - s) / d; }
__device__ __forceinline__ double2 ic_fin(double2 b, double2 s, double2 d) { return zdiv(zsub(b, s), d); }
__device__ __forceinline__ void ic_st(double* p, double v) { __stcg(p, v); }
__device__ __forceinline__ void ic_st(double2* p, double2 v) { __stcg(p, v); }
__device__ __forceinline__ void ic_st(ZF* p, double2 v) { __stcg(reinterpret_cast<float2*>(p), make_float2((float)v.x, (float)v.y)); }
__device__ __forceinline__ double ic_zero(double) { return 0.0; }
__device__ __forceinline__ double2 ic_zero(double2) { return make_double2(0.0, 0.0); }

template <class T, bool UPPER>
__global__ void __launch_bounds__(256) k_sptrsv(IcDev F, const T* __restrict__ b, T* x, DevState* st)
{
	pdl_enter();
	if (st_done(st)) return;
	const T* val = static_cast<const T*>(F.val);
	const int epoch = F.ctl->epoch + 1;
	const int lane = threadIdx.x & 31;
	while (true)
	{
		unsigned int base = 0;
		if (lane == 0) base = atomicAdd(&F.ctl->next, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= (unsigned int)F.n_pos) break;
		const int row = F.order[base + lane];
		if (row < 0) continue;
		const int kb = F.rp[row], ke = F.rp[row + 1];
		const int kd = UPPER ? kb : ke - 1;          // the diagonal entry
		auto sum = ic_zero(ic_ld(val + kd));
		for (int k = UPPER ? kb + 1 : kb; k < (UPPER ? ke : ke - 1); k++)
		{
			const int c = F.ci[k];
			while (ld_acquire_gpu(F.ready + c) != epoch) { }
			sum = ic_fma(ic_ld(val + k), ic_ld(x + c), sum);
		}
		ic_st(x + row, ic_fin(ic_ld(b + row), sum, ic_ld(val + kd)));
		__threadfence();
		st_release_gpu(F.ready + row, epoch);
	}
	__syncthreads();
	if (threadIdx.x == 0 && atomicAdd(&F.ctl->ticket, 1u) == gridDim.x - 1)
	{	// the block that leaves last opens the next solve: new epoch, counter back to the first position
		F.ctl->epoch = epoch; F.ctl->next = 0u; F.ctl->ticket = 0u;
		__threadfence();
	}
}

template <class T, bool UPPER>
inline void launch_sptrsv(const IcDev& F, const T* b, T* x, DevState* st, cudaStream_t s)
{
	int blocks = (F.n_pos + 255) / 256;
	if (blocks > 148 * 8) blocks = 148 * 8;
	if (blocks < 1) blocks = 1;
	launch_k(k_sptrsv<T, UPPER>, blocks, 256, 0, s, F, b, x, st);
}

}  // namespace lcgb200
