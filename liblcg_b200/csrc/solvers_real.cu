// solvers_real.cu — the real-valued iteration loops (CG, PCG, CGS, BICGSTAB, BICGSTAB2, PG, SPG) as sequences
// of fused kernels.  Each step functor cites the reference lines whose arithmetic it carries.
//
// HBM passes per iteration (vector passes outside the SpMV's own x/y):
//   CG  6 + 3 | PCG 8 + 3 | CGS 1 + 4 + 7 + 5 | BICGSTAB 1 + 3 + 1 + 7 + 4 | PG 3 + 7
#include "solvers.cuh"

// The file is compiled twice: as it is (solve_real, the fast path), and with -DLCG_REFORDER -fmad=false (solve_real_x: the
// reference-order build, common.cuh "arithmetic variants" / exact.cuh).  The step functors below are written so that each
// expression has the reference's shape — fma(a, b, c) stands where the reference computes a*b + c — which makes the second
// build bit-identical to the reference's CPU loops.
#ifdef LCG_REFORDER
#define LCG_REAL_NS rx64
#else
#define LCG_REAL_NS r64
#endif

namespace lcgb200 {
namespace LCG_REAL_NS {

__device__ __forceinline__ double box(double lo, double hi, double a)
{	// lcg_set2box with closed bounds (algebra.cpp:50-58 as called at lcg.cpp:1089)
	if (a >= hi) return hi;
	if (a <= lo) return lo;
	return a;
}

// ======================================================================================== shared epilogues
// alpha = rho / (w . A x); w = x when w == nullptr.  CG lcg.cpp:234-235, PCG :389-390, CGS :548-553, BICGSTAB :720-725
struct EpiDotAlpha {
	static constexpr int NRED = 1;
	static constexpr bool ACTIVE = true;
	const double* w;
	__device__ void begin(const DevState*) {}
	__device__ void row(int i, double yi, double xi, double* acc) const { acc[0] = fma(w ? w[i] : xi, yi, acc[0]); }
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_ALPHA] = st->sc[SC_RHO] / tot[0]; }
};

// omega = (As . s) / (As . As).  lcg.cpp:735-741
struct EpiOmega {
	static constexpr int NRED = 2;
	static constexpr bool ACTIVE = true;
	__device__ void begin(const DevState*) {}
	__device__ void row(int, double yi, double xi, double* acc) const { acc[0] = fma(yi, xi, acc[0]); acc[1] = fma(yi, yi, acc[1]); }
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_OMEGA] = tot[0] / tot[1]; }
};

// SPG objective of the trial point: q = sum(0.5 m_new A m_new - B m_new).  lcg.cpp:1359-1363
struct EpiSpgQ {
	static constexpr int NRED = 1;
	static constexpr bool ACTIVE = true;
	const double* B;
	__device__ void begin(const DevState*) {}
	__device__ void row(int i, double yi, double xi, double* acc) const { acc[0] += (0.5 * xi * yi - B[i] * xi); }
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_QK] = tot[0]; }
};

// ======================================================================================== CG  (lcg.cpp:143-274)
struct OpCgInit : OpBase {	// g = Ad - B, d = -g, m.m, g.g  (lcg.cpp:171-183) + first loop head
	static constexpr int NRED = 2, W = 2;
	const double* m; const double* Ad; const double* B; double* g; double* d;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), va = dv_load<V>(Ad + i), vb = dv_load<V>(B + i), vg, vd;
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vg.v[k] = va.v[k] - vb.v[k]; vd.v[k] = -1.0 * vg.v[k];
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vg.v[k], vg.v[k], acc[1]);
		}
		dv_store<V>(g + i, vg); dv_store<V>(d + i, vd);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RHO] = tot[1]; st->sc[SC_RMOD] = tot[1];
		first_head_real(st, tot[1], mm);
	}
};

struct OpCgUpdate : OpBase {	// m += a d, g += a Ad, m.m, NaN, g.g, beta (lcg.cpp:238-257) + next loop head
	static constexpr int NRED = 2, W = 2;
	double* m; const double* d; double* g; const double* Ad; double ak;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vd = dv_load<V>(d + i), vg = dv_load<V>(g + i), va = dv_load<V>(Ad + i);
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vm.v[k] = fma(ak, vd.v[k], vm.v[k]); vg.v[k] = fma(ak, va.v[k], vg.v[k]);
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vg.v[k], vg.v[k], acc[1]);
		}
		dv_store<V>(m + i, vm); dv_store<V>(g + i, vg);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (tot[0] != tot[0]) { st->ret = RC_NAN; st->done = 1; return; }	// a NaN in m makes m.m NaN
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm;
		st->sc[SC_BETA] = tot[1] / st->sc[SC_RHO];
		st->sc[SC_RHO] = tot[1]; st->sc[SC_RMOD] = tot[1];
		loop_head_real(st, tot[1], mm);
	}
};

struct OpCgDir : OpBase {	// d = beta d - g (lcg.cpp:260-263)
	static constexpr int NRED = 0, W = 2;
	double* d; const double* g; double bk;
	__device__ void begin(const DevState* st) { bk = st->sc[SC_BETA]; }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vd = dv_load<V>(d + i), vg = dv_load<V>(g + i);
#pragma unroll
		for (int k = 0; k < V; k++) vd.v[k] = bk * vd.v[k] - vg.v[k];
		dv_store<V>(d + i, vd);
	}
};

static int run_cg(Engine& E, const Operator<double>& A, double* m, const double* B, size_t n, size_t next)
{
	double* g = E.alloc<double>(next); double* d = E.alloc<double>(next); double* Ad = E.alloc<double>(next);
	E.spmv(A, m, Ad, EpiNone<double>{});
	E.vec_push(OpCgInit{{}, m, Ad, B, g, d}, n, d);
	std::function<void(int)> batch;
	if (E.small_system(A)) batch = [&](int k) { E.fused(k, 1, E.ph_spmv(A, d, Ad, EpiDotAlpha{nullptr}), E.ph_vec(OpCgUpdate{{}, m, d, g, Ad, 0.0}, n), E.ph_vec(OpCgDir{{}, d, g, 0.0}, n)); };
	return E.run([&]() {
		E.spmv(A, d, Ad, EpiDotAlpha{nullptr});
		E.vec2_push(OpCgUpdate{{}, m, d, g, Ad, 0.0}, OpCgDir{{}, d, g, 0.0}, n, d);
		return false;
	}, batch);
}

// ======================================================================================== PCG (lcg.cpp:293-434)
// JAC = built-in Jacobi z = r / diag fused into the update (the reference does it with a separate
// lcg_vecDvecD_element_wise kernel inside the user's Mfp, sample10.cu:117 / algebra_cuda.cu:69-77).
template <bool JAC>
struct OpPcgInit : OpBase {	// r = B - Ad [, z = r/diag, d = z], m.m, r.r [, z.r]  (lcg.cpp:316-338)
	static constexpr int NRED = 3, W = 2;
	const double* m; const double* Ad; const double* B; const double* diag; double* r; double* z; double* d;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), va = dv_load<V>(Ad + i), vb = dv_load<V>(B + i), vr, vz;
		DV<V> vdg; if (JAC) vdg = dv_load<V>(diag + i);
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vr.v[k] = vb.v[k] - va.v[k];
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vr.v[k], vr.v[k], acc[1]);
			if (JAC) { vz.v[k] = vr.v[k] / vdg.v[k]; acc[2] = fma(vz.v[k], vr.v[k], acc[2]); }
		}
		dv_store<V>(r + i, vr);
		if (JAC) { dv_store<V>(z + i, vz); dv_store<V>(d + i, vz); }
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (!JAC) { st->sc[SC_PART0] = tot[0]; st->sc[SC_PART1] = tot[1]; return; }
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = tot[1]; st->sc[SC_RHO] = tot[2];
		first_head_real(st, tot[1], mm);
	}
};

struct OpPcgInitZ : OpBase {	// user Mfp path: d = z, z.r, then the head (lcg.cpp:324-338)
	static constexpr int NRED = 1, W = 2;
	const double* z; const double* r; double* d;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vz = dv_load<V>(z + i), vr = dv_load<V>(r + i);
#pragma unroll
		for (int k = 0; k < V; k++) acc[0] = fma(vz.v[k], vr.v[k], acc[0]);
		dv_store<V>(d + i, vz);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		double mm = st->sc[SC_PART0] < 1.0 ? 1.0 : st->sc[SC_PART0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = st->sc[SC_PART1]; st->sc[SC_RHO] = tot[0];
		first_head_real(st, st->sc[SC_PART1], mm);
	}
};

__device__ __forceinline__ void pcg_tail(DevState* st, double mm_raw, double rr, double zr)
{	// lcg.cpp:401-416 + next loop head
	if (mm_raw != mm_raw) { st->ret = RC_NAN; st->done = 1; return; }
	double mm = mm_raw < 1.0 ? 1.0 : mm_raw;
	st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = rr;
	st->sc[SC_BETA] = zr / st->sc[SC_RHO];
	st->sc[SC_RHO] = zr;
	loop_head_real(st, rr, mm);
}

template <bool JAC>
struct OpPcgUpdate : OpBase {	// m += a d, r -= a Ad [, z = r/diag], m.m, r.r [, z.r]  (lcg.cpp:392-414)
	static constexpr int NRED = 3, W = 2;
	double* m; const double* d; double* r; const double* Ad; const double* diag; double* z; double ak;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vd = dv_load<V>(d + i), vr = dv_load<V>(r + i), va = dv_load<V>(Ad + i), vz;
		DV<V> vdg; if (JAC) vdg = dv_load<V>(diag + i);
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vm.v[k] = fma(ak, vd.v[k], vm.v[k]); vr.v[k] = fma(-ak, va.v[k], vr.v[k]);
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vr.v[k], vr.v[k], acc[1]);
			if (JAC) { vz.v[k] = vr.v[k] / vdg.v[k]; acc[2] = fma(vz.v[k], vr.v[k], acc[2]); }
		}
		dv_store<V>(m + i, vm); dv_store<V>(r + i, vr);
		if (JAC) dv_store<V>(z + i, vz);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (!JAC) { st->sc[SC_PART0] = tot[0]; st->sc[SC_PART1] = tot[1]; return; }
		pcg_tail(st, tot[0], tot[1], tot[2]);
	}
};

struct OpPcgZr : OpBase {	// user Mfp path: z.r after the callback
	static constexpr int NRED = 1, W = 2;
	const double* z; const double* r;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vz = dv_load<V>(z + i), vr = dv_load<V>(r + i);
#pragma unroll
		for (int k = 0; k < V; k++) acc[0] = fma(vz.v[k], vr.v[k], acc[0]);
	}
	__device__ void finish(DevState* st, const double* tot) const { pcg_tail(st, st->sc[SC_PART0], st->sc[SC_PART1], tot[0]); }
};

struct OpPcgDir : OpBase {	// d = z + beta d (lcg.cpp:418-422)
	static constexpr int NRED = 0, W = 2;
	double* d; const double* z; double bk;
	__device__ void begin(const DevState* st) { bk = st->sc[SC_BETA]; }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vd = dv_load<V>(d + i), vz = dv_load<V>(z + i);
#pragma unroll
		for (int k = 0; k < V; k++) vd.v[k] = fma(bk, vd.v[k], vz.v[k]);
		dv_store<V>(d + i, vd);
	}
};

static int run_pcg(Engine& E, const Operator<double>& A, double* m, const double* B, size_t n, size_t next)
{
	double* r = E.alloc<double>(next); double* z = E.alloc<double>(next);
	double* d = E.alloc<double>(next); double* Ad = E.alloc<double>(next);
	const bool jac = (A.diag != nullptr);
	E.spmv(A, m, Ad, EpiNone<double>{});
	if (jac) E.vec_push(OpPcgInit<true>{{}, m, Ad, B, A.diag, r, z, d}, n, d);
	else
	{
		E.vec(OpPcgInit<false>{{}, m, Ad, B, nullptr, r, z, d}, n);
		E.precondition(A, r, z);
		E.vec_push(OpPcgInitZ{{}, z, r, d}, n, d);
	}
	std::function<void(int)> batch;
	if (jac && E.small_system(A))
		batch = [&](int k) { E.fused(k, 1, E.ph_spmv(A, d, Ad, EpiDotAlpha{nullptr}), E.ph_vec(OpPcgUpdate<true>{{}, m, d, r, Ad, A.diag, z, 0.0}, n), E.ph_vec(OpPcgDir{{}, d, z, 0.0}, n)); };
	return E.run([&]() {
		E.spmv(A, d, Ad, EpiDotAlpha{nullptr});
		if (jac) E.vec2_push(OpPcgUpdate<true>{{}, m, d, r, Ad, A.diag, z, 0.0}, OpPcgDir{{}, d, z, 0.0}, n, d);
		else
		{
			E.vec(OpPcgUpdate<false>{{}, m, d, r, Ad, nullptr, z, 0.0}, n);
			E.precondition(A, r, z);
			E.vec2_push(OpPcgZr{{}, z, r}, OpPcgDir{{}, d, z, 0.0}, n, d);
		}
		return false;
	}, batch);
}

// ======================================================================================== CGS (lcg.cpp:437-612)
struct OpResInit : OpBase {	// p = [u =] r0 = r = B - Ax; r.r0, m.m, r.r (lcg.cpp:480-497, 652-669)
	static constexpr int NRED = 2, W = 2;
	const double* m; const double* Ax; const double* B; double* r; double* r0; double* p; double* u;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), va = dv_load<V>(Ax + i), vb = dv_load<V>(B + i), vr;
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vr.v[k] = vb.v[k] - va.v[k];
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vr.v[k], vr.v[k], acc[1]);
		}
		dv_store<V>(r + i, vr); dv_store<V>(r0 + i, vr); dv_store<V>(p + i, vr);
		if (u) dv_store<V>(u + i, vr);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = tot[1]; st->sc[SC_RHO] = tot[1];	// r.r0 == r.r at start
		first_head_real(st, tot[1], mm);
	}
};

struct OpCgsQW : OpBase {	// q = u - a Ap, w = u + q (lcg.cpp:555-560)
	static constexpr int NRED = 0, W = 2;
	const double* u; const double* Ax; double* q; double* w; double ak;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vu = dv_load<V>(u + i), va = dv_load<V>(Ax + i), vq, vw;
#pragma unroll
		for (int k = 0; k < V; k++) { vq.v[k] = fma(-ak, va.v[k], vu.v[k]); vw.v[k] = vu.v[k] + vq.v[k]; }
		dv_store<V>(q + i, vq); dv_store<V>(w + i, vw);
	}
};

struct OpCgsUpdate : OpBase {	// m += a w, r -= a Aw; m.m, r.r, r.r0; beta (lcg.cpp:564-590) + head
	static constexpr int NRED = 3, W = 2;
	double* m; const double* w; double* r; const double* Ax; const double* r0; double ak;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vw = dv_load<V>(w + i), vr = dv_load<V>(r + i), va = dv_load<V>(Ax + i), v0 = dv_load<V>(r0 + i);
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vm.v[k] = fma(ak, vw.v[k], vm.v[k]); vr.v[k] = fma(-ak, va.v[k], vr.v[k]);
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vr.v[k], vr.v[k], acc[1]); acc[2] = fma(vr.v[k], v0.v[k], acc[2]);
		}
		dv_store<V>(m + i, vm); dv_store<V>(r + i, vr);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (tot[0] != tot[0]) { st->ret = RC_NAN; st->done = 1; return; }
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = tot[1];
		st->sc[SC_BETA] = tot[2] / st->sc[SC_RHO];
		st->sc[SC_RHO] = tot[2];
		loop_head_real(st, tot[1], mm);
	}
};

struct OpCgsDir : OpBase {	// u = r + b q, p = u + b (q + b p) (lcg.cpp:592-597)
	static constexpr int NRED = 0, W = 2;
	const double* r; const double* q; double* u; double* p; double bk;
	__device__ void begin(const DevState* st) { bk = st->sc[SC_BETA]; }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vr = dv_load<V>(r + i), vq = dv_load<V>(q + i), vp = dv_load<V>(p + i), vu;
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vu.v[k] = fma(bk, vq.v[k], vr.v[k]);
			vp.v[k] = fma(bk, fma(bk, vp.v[k], vq.v[k]), vu.v[k]);
		}
		dv_store<V>(u + i, vu); dv_store<V>(p + i, vp);
	}
};

static int run_cgs(Engine& E, const Operator<double>& A, double* m, const double* B, size_t n, size_t next)
{
	double* r = E.alloc<double>(next); double* r0 = E.alloc<double>(next); double* p = E.alloc<double>(next);
	double* Ax = E.alloc<double>(next); double* u = E.alloc<double>(next); double* q = E.alloc<double>(next);
	double* w = E.alloc<double>(next);
	E.spmv(A, m, Ax, EpiNone<double>{});
	E.vec_push(OpResInit{{}, m, Ax, B, r, r0, p, u}, n, p);
	std::function<void(int)> batch;
	if (E.small_system(A))
		batch = [&](int k) { E.fused(k, 2, E.ph_spmv(A, p, Ax, EpiDotAlpha{r0}), E.ph_vec(OpCgsQW{{}, u, Ax, q, w, 0.0}, n), E.ph_spmv(A, w, Ax, EpiNone<double>{}),
			E.ph_vec(OpCgsUpdate{{}, m, w, r, Ax, r0, 0.0}, n), E.ph_vec(OpCgsDir{{}, r, q, u, p, 0.0}, n)); };
	return E.run([&]() {
		E.spmv(A, p, Ax, EpiDotAlpha{r0});
		E.vec_push(OpCgsQW{{}, u, Ax, q, w, 0.0}, n, w);
		E.spmv(A, w, Ax, EpiNone<double>{});
		E.vec2_push(OpCgsUpdate{{}, m, w, r, Ax, r0, 0.0}, OpCgsDir{{}, r, q, u, p, 0.0}, n, p);
		return false;
	}, batch);
}

// ================================================================ BICGSTAB (lcg.cpp:629-794) / BICGSTAB2 (:812-1034)
template <bool HALF>
struct OpBicgS : OpBase {	// s = r - a Ap (lcg.cpp:727-731); HALF: + s.s and the half-step head (lcg.cpp:918-950)
	static constexpr int NRED = HALF ? 1 : 0, W = 2;
	const double* r; const double* Ap; double* s; double ak;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vr = dv_load<V>(r + i), va = dv_load<V>(Ap + i), vs;
#pragma unroll
		for (int k = 0; k < V; k++) { vs.v[k] = fma(-ak, va.v[k], vr.v[k]); if (HALF) acc[0] = fma(vs.v[k], vs.v[k], acc[0]); }
		dv_store<V>(s + i, vs);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (!HALF) return;
		const double residual = sqrt(tot[0]) / (double)st->n_global;
		st->residual = residual; st->k_report = st->t; st->checks++;
		if (residual <= st->eps) { st->flag = 1; return; }	// OpBicgHalf applies m += a p and ends the solve
		if (st->max_it > 0 && st->t + 1 > st->max_it) { st->ret = RC_MAXIT; st->done = 1; return; }
		st->t++;
	}
};

struct OpBicgHalf : OpBase {	// converged on the half step: m += a p, NaN check (lcg.cpp:930-941)
	static constexpr int NRED = 1, W = 2;
	double* m; const double* p; double ak;
	__device__ bool active(const DevState* st) const { return *((volatile const int*)&st->flag) != 0; }
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vp = dv_load<V>(p + i);
#pragma unroll
		for (int k = 0; k < V; k++) { vm.v[k] = fma(ak, vp.v[k], vm.v[k]); acc[0] = fma(vm.v[k], vm.v[k], acc[0]); }
		dv_store<V>(m + i, vm);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		st->ret = (tot[0] != tot[0]) ? RC_NAN : RC_CONVERGENCE; st->done = 1;
	}
};

template <bool RESTART>
struct OpBicgUpdate : OpBase {	// m += a p + w s, r = s - w As; m.m, r.r, r.r0; beta / restart (lcg.cpp:743-774, 962-1013)
	static constexpr int NRED = 3, W = 2;
	double* m; const double* p; const double* s; const double* As; double* r; const double* r0; double ak, wk;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_ALPHA]; wk = st->sc[SC_OMEGA]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vp = dv_load<V>(p + i), vs = dv_load<V>(s + i), va = dv_load<V>(As + i), v0 = dv_load<V>(r0 + i), vr;
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vm.v[k] += fma(ak, vp.v[k], wk * vs.v[k]);
			vr.v[k] = fma(-wk, va.v[k], vs.v[k]);
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vr.v[k], vr.v[k], acc[1]); acc[2] = fma(vr.v[k], v0.v[k], acc[2]);
		}
		dv_store<V>(m + i, vm); dv_store<V>(r + i, vr);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (tot[0] != tot[0]) { st->ret = RC_NAN; st->done = 1; return; }
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = tot[1];
		if (RESTART && fabs(tot[2]) < st->restart_eps)
		{	// r0 = p = r (done by OpBicgDir); r.r0 becomes r.r (lcg.cpp:993-1009)
			st->half = 1; st->sc[SC_RHO] = tot[1];
		}
		else
		{
			st->half = 0;
			st->sc[SC_BETA] = (st->sc[SC_ALPHA] / st->sc[SC_OMEGA]) * tot[2] / st->sc[SC_RHO];
			st->sc[SC_RHO] = tot[2];
		}
		loop_head_real(st, tot[1], mm);
	}
};

template <bool RESTART>
struct OpBicgDir : OpBase {	// p = r + b (p - w Ap) (lcg.cpp:776-780), or the restart r0 = p = r (lcg.cpp:996-1000)
	static constexpr int NRED = 0, W = 2;
	const double* r; double* p; const double* Ap; double* r0; double bk, wk; int restart;
	__device__ void begin(const DevState* st) { bk = st->sc[SC_BETA]; wk = st->sc[SC_OMEGA]; restart = RESTART ? st->half : 0; }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vr = dv_load<V>(r + i);
		if (RESTART && restart) { dv_store<V>(r0 + i, vr); dv_store<V>(p + i, vr); return; }
		DV<V> vp = dv_load<V>(p + i), va = dv_load<V>(Ap + i);
#pragma unroll
		for (int k = 0; k < V; k++) vp.v[k] = fma(bk, fma(-wk, va.v[k], vp.v[k]), vr.v[k]);
		dv_store<V>(p + i, vp);
	}
};

template <bool RESTART>
static int run_bicgstab(Engine& E, const Operator<double>& A, double* m, const double* B, size_t n, size_t next, bool abs_diff)
{
	double* r = E.alloc<double>(next); double* r0 = E.alloc<double>(next); double* p = E.alloc<double>(next);
	double* Ax = E.alloc<double>(next); double* s = E.alloc<double>(next); double* Ap = E.alloc<double>(next);
	const bool half = RESTART && abs_diff;
	if (half) E.capturable = false;	// Pfp may be consulted in the middle of the iteration
	E.spmv(A, m, Ax, EpiNone<double>{});
	E.vec_push(OpResInit{{}, m, Ax, B, r, r0, p, nullptr}, n, p);
	std::function<void(int)> batch;
	if (!half && E.small_system(A))
		batch = [&](int k) { E.fused(k, 2, E.ph_spmv(A, p, Ap, EpiDotAlpha{r0}), E.ph_vec(OpBicgS<false>{{}, r, Ap, s, 0.0}, n), E.ph_spmv(A, s, Ax, EpiOmega{}),
			E.ph_vec(OpBicgUpdate<RESTART>{{}, m, p, s, Ax, r, r0, 0.0, 0.0}, n), E.ph_vec(OpBicgDir<RESTART>{{}, r, p, Ap, r0, 0.0, 0.0, 0}, n)); };
	return E.run([&]() {
		E.spmv(A, p, Ap, EpiDotAlpha{r0});
		if (half)
		{
			E.vec_push(OpBicgS<true>{{}, r, Ap, s, 0.0}, n, s);
			if (E.sync_point()) return true;	// Pfp sees m before the half update, as in the reference
			E.vec(OpBicgHalf{{}, m, p, 0.0}, n);
		}
		else E.vec_push(OpBicgS<false>{{}, r, Ap, s, 0.0}, n, s);
		E.spmv(A, s, Ax, EpiOmega{});
		E.vec2_push(OpBicgUpdate<RESTART>{{}, m, p, s, Ax, r, r0, 0.0, 0.0}, OpBicgDir<RESTART>{{}, r, p, Ap, r0, 0.0, 0.0, 0}, n, p);
		return false;
	}, batch);
}

// ======================================================================================== PG (lcg.cpp:1054-1204)
struct OpBox : OpBase {	// m = P_box(m) (lcg.cpp:1086-1090)
	static constexpr int NRED = 0, W = 2;
	double* m; const double* lo; const double* hi;
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vm = dv_load<V>(m + i), vl = dv_load<V>(lo + i), vh = dv_load<V>(hi + i);
#pragma unroll
		for (int k = 0; k < V; k++) vm.v[k] = box(vl.v[k], vh.v[k], vm.v[k]);
		dv_store<V>(m + i, vm);
	}
};

template <bool SPG>
struct OpPgInit : OpBase {	// g = Ad - B; m.m, g.g [, q0 = sum(0.5 m Ad - B m)] (lcg.cpp:1094-1105, 1304-1309)
	static constexpr int NRED = SPG ? 3 : 2, W = 2;
	const double* m; const double* Ad; const double* B; double* g;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), va = dv_load<V>(Ad + i), vb = dv_load<V>(B + i), vg;
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			vg.v[k] = va.v[k] - vb.v[k];
			acc[0] = fma(vm.v[k], vm.v[k], acc[0]); acc[1] = fma(vg.v[k], vg.v[k], acc[1]);
			if (SPG) acc[2] += (0.5 * vm.v[k] * va.v[k] - vb.v[k] * vm.v[k]);
		}
		dv_store<V>(g + i, vg);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		double mm = tot[0] < 1.0 ? 1.0 : tot[0];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = tot[1];
		if (SPG) st->sc[SC_QK] = tot[2];
		first_head_real(st, tot[1], mm);
	}
};

struct OpPgStep : OpBase {	// m_new = P_box(m - a g) (lcg.cpp:1155-1159)
	static constexpr int NRED = 0, W = 2;
	const double* m; const double* g; const double* lo; const double* hi; double* mn; double ak;
	__device__ void begin(const DevState* st) { ak = st->sc[SC_STEP]; }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vm = dv_load<V>(m + i), vg = dv_load<V>(g + i), vl = dv_load<V>(lo + i), vh = dv_load<V>(hi + i), vn;
#pragma unroll
		for (int k = 0; k < V; k++) vn.v[k] = box(vl.v[k], vh.v[k], fma(-ak, vg.v[k], vm.v[k]));
		dv_store<V>(mn + i, vn);
	}
};

struct OpPgUpdate : OpBase {	// g_new = Ad - B, s = m_new - m, y = g_new - g; s.s, s.y; m = m_new, g = g_new; m.m, g.g
	static constexpr int NRED = 4, W = 2;	// (lcg.cpp:1163-1190, 1404-1431) + BB step + head
	double* m; double* g; const double* mn; const double* Ad; const double* B;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vg = dv_load<V>(g + i), vn = dv_load<V>(mn + i), va = dv_load<V>(Ad + i), vb = dv_load<V>(B + i), gn;
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			gn.v[k] = va.v[k] - vb.v[k];
			const double s = vn.v[k] - vm.v[k], y = gn.v[k] - vg.v[k];
			acc[0] = fma(s, s, acc[0]); acc[1] = fma(s, y, acc[1]);
			acc[2] = fma(vn.v[k], vn.v[k], acc[2]); acc[3] = fma(gn.v[k], gn.v[k], acc[3]);
		}
		dv_store<V>(m + i, vn); dv_store<V>(g + i, gn);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		st->sc[SC_STEP] = tot[0] / tot[1];
		double mm = tot[2] < 1.0 ? 1.0 : tot[2];
		st->sc[SC_MMOD] = mm; st->sc[SC_RMOD] = tot[3];
		loop_head_real(st, tot[3], mm);
	}
};

static int run_pg(Engine& E, const Operator<double>& A, double* m, const double* B, const double* lo, const double* hi, size_t n, size_t next)
{
	double* g = E.alloc<double>(next); double* Ad = E.alloc<double>(next); double* mn = E.alloc<double>(next);
	E.vec_push(OpBox{{}, m, lo, hi}, n, m);
	E.spmv(A, m, Ad, EpiNone<double>{});
	E.vec(OpPgInit<false>{{}, m, Ad, B, g}, n);
	std::function<void(int)> batch;
	if (E.small_system(A))
		batch = [&](int k) { E.fused(k, 1, E.ph_vec(OpPgStep{{}, m, g, lo, hi, mn, 0.0}, n), E.ph_spmv(A, mn, Ad, EpiNone<double>{}), E.ph_vec(OpPgUpdate{{}, m, g, mn, Ad, B}, n)); };
	return E.run([&]() {
		E.vec_push(OpPgStep{{}, m, g, lo, hi, mn, 0.0}, n, mn);
		E.spmv(A, mn, Ad, EpiNone<double>{});
		E.vec(OpPgUpdate{{}, m, g, mn, Ad, B}, n);
		return false;
	}, batch);
}

// ======================================================================================== SPG (lcg.cpp:1224-1447)
struct OpSpgDir : OpBase {	// d = P_box(m - l g) - m; g.d (lcg.cpp:1345-1349, 1365-1369)
	static constexpr int NRED = 1, W = 2;
	const double* m; const double* g; const double* lo; const double* hi; double* d; double lk;
	__device__ void begin(const DevState* st) { lk = st->sc[SC_STEP]; }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		DV<V> vm = dv_load<V>(m + i), vg = dv_load<V>(g + i), vl = dv_load<V>(lo + i), vh = dv_load<V>(hi + i), vd;
#pragma unroll
		for (int k = 0; k < V; k++) { vd.v[k] = box(vl.v[k], vh.v[k], fma(-lk, vg.v[k], vm.v[k])) - vm.v[k]; acc[0] = fma(vg.v[k], vd.v[k], acc[0]); }
		dv_store<V>(d + i, vd);
	}
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_GD] = tot[0]; }
};

struct OpSpgTrial : OpBase {	// m_new = m + a d (lcg.cpp:1352-1355, 1381-1384); a comes from the host line search
	static constexpr int NRED = 0, W = 2;
	const double* m; const double* d; double* mn; double alpha;
	template <int V> __device__ void elem(size_t i, double*) const
	{
		DV<V> vm = dv_load<V>(m + i), vd = dv_load<V>(d + i), vn;
#pragma unroll
		for (int k = 0; k < V; k++) vn.v[k] = fma(alpha, vd.v[k], vm.v[k]);
		dv_store<V>(mn + i, vn);
	}
};

static int run_spg(Engine& E, const Operator<double>& A, double* m, const double* B, const double* lo, const double* hi,
	size_t n, size_t next, const lcgb200_para& para)
{
	double* g = E.alloc<double>(next); double* Ad = E.alloc<double>(next); double* mn = E.alloc<double>(next);
	double* d = E.alloc<double>(next);
	E.capturable = false;	// host-side line search
	E.vec_push(OpBox{{}, m, lo, hi}, n, m);
	E.spmv(A, m, Ad, EpiNone<double>{});
	E.vec(OpPgInit<true>{{}, m, Ad, B, g}, n);
	// the non-monotone history lives on the host: the line search is data dependent (lcg.cpp:1377-1399)
	std::vector<double> qm((size_t)para.maxi_m, -1e+30);
	bool have_q0 = false;
	return E.run([&]() {
		if (!have_q0)
		{	// first pass: fetch q0 (and let the "already optimised" / first head be seen)
			if (E.sync_always()) return true;
			qm[0] = E.h_st->sc[SC_QK]; have_q0 = true;
		}
		E.vec(OpSpgDir{{}, m, g, lo, hi, d, 0.0}, n);
		double alpha = 1.0;
		int t_now = 0;
		while (true)
		{
			E.vec_push(OpSpgTrial{{}, m, d, mn, alpha}, n, mn);
			E.spmv(A, mn, Ad, EpiSpgQ{B});
			E.read_state();
			// the loop head of the previous iteration (read here for the first time when no progress callback forces a round
			// trip of its own) may have ended the solve: everything enqueued since then was a no-op
			if (E.h_st->done) return E.sync_always();
			const double qk = E.h_st->sc[SC_QK], gd = E.h_st->sc[SC_GD];
			t_now = E.h_st->t;
			const double amod = para.sigma * alpha * gd;
			double qmax = qm[0];
			for (int i = 1; i < para.maxi_m; i++) qmax = (qmax >= qm[(size_t)i]) ? qmax : qm[(size_t)i];
			if (!(qk > qmax + amod)) { qm[(size_t)((t_now + 1) % para.maxi_m)] = qk; break; }
			alpha = alpha * para.beta;
			if (!(alpha > 0.0) || alpha < 1e-300) { qm[(size_t)((t_now + 1) % para.maxi_m)] = qk; break; }	// guard: the reference would spin forever
		}
		E.vec(OpPgUpdate{{}, m, g, mn, Ad, B}, n);
		// one host round trip per line-search trial is inherent (the search is data dependent); the loop head needs a second
		// one only when a progress callback must see it — otherwise the next iteration's first trial reads it
		return E.pf ? E.sync_always() : false;
	});
}

// ======================================================================================== dispatch
static int dispatch(Engine& E, const Operator<double>& A, int solver_id, double* m, const double* B, const double* lo, const double* hi,
	const lcgb200_para& para, size_t n, size_t next)
{
	switch (solver_id)
	{
		case LCGB200_CG: return run_cg(E, A, m, B, n, next);
		case LCGB200_PCG: return run_pcg(E, A, m, B, n, next);
		case LCGB200_BICGSTAB: return run_bicgstab<false>(E, A, m, B, n, next, para.abs_diff != 0);
		case LCGB200_BICGSTAB2: return run_bicgstab<true>(E, A, m, B, n, next, para.abs_diff != 0);
		case LCGB200_PG: return run_pg(E, A, m, B, lo, hi, n, next);
		case LCGB200_SPG: return run_spg(E, A, m, B, lo, hi, n, next, para);
		case LCGB200_CGS: default: return run_cgs(E, A, m, B, n, next);
	}
}

}  // namespace LCG_REAL_NS

#ifdef LCG_REFORDER
int solve_real_x(Engine& E, const Operator<double>& A, int solver_id, double* m, const double* B, const double* lo, const double* hi,
	const lcgb200_para& para, size_t n, size_t next)
{
	return rx64::dispatch(E, A, solver_id, m, B, lo, hi, para, n, next);
}
#else
int solve_real(Engine& E, const Operator<double>& A, int solver_id, double* m, const double* B, const double* lo, const double* hi,
	const lcgb200_para& para, size_t n, size_t next)
{
	return r64::dispatch(E, A, solver_id, m, B, lo, hi, para, n, next);
}

int real_vector_count(int solver_id)
{
	switch (solver_id)
	{
		case LCGB200_CG: return 3;
		case LCGB200_PCG: return 4;
		case LCGB200_BICGSTAB: case LCGB200_BICGSTAB2: return 6;
		case LCGB200_PG: return 3;
		case LCGB200_SPG: return 4;
		default: return 7;
	}
}

#endif

}  // namespace lcgb200
