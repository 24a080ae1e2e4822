// solvers_complex.cu — the complex iteration loops: BiCG, BiCG for symmetric A, CGS, BiCGSTAB, TFQMR
// (reference CPU versions clcg.cpp:77-882) and Jacobi/user-preconditioned PCG (clcg_cuda.cu:403-559,
// clcg_eigen.cpp:577-683).  Vectors are cuDoubleComplex == double2 (one 128-bit access per element).
//
// Inner-product conventions (lcg_complex.cpp:143-167): clcg_inner = sum conj(a_i) b_i ("conj-first"),
// clcg_dot = sum a_i b_i (unconjugated, used by BICG_SYM and PCG).
#include "solvers.cuh"
#include <cstdlib>
#include <ctime>

// The file is compiled twice: as it is for cuDoubleComplex vectors (solve_complex), and with -DLCG_CPLX_FLOAT for cuComplex
// vectors (solve_complexf: the entry points of clcg_cudaf.h:81-105, reference implementation clcg_cudaf.cu).  ZV is the type
// of a vector / matrix element in memory; the arithmetic of every step runs on Z = double2 either way.
// A third build (-DLCG_REFORDER -fmad=false: solve_complex_x) is the reference-order variant of the double-precision loops
// (common.cuh "arithmetic variants" / exact.cuh).
#if defined(LCG_CPLX_FLOAT)
#define LCG_CPLX_NS cf32
#elif defined(LCG_REFORDER)
#define LCG_CPLX_NS cx64
#else
#define LCG_CPLX_NS cf64
#endif

namespace lcgb200 {
namespace LCG_CPLX_NS {

typedef double2 Z;
#ifdef LCG_CPLX_FLOAT
typedef ZF ZV;
#else
typedef double2 ZV;
#endif

__device__ __forceinline__ void acc_inner(double* acc, Z a, Z b)	// acc += conj(a) b
{
	acc[0] += (a.x * b.x + a.y * b.y);
	acc[1] += (a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ void acc_dotu(double* acc, Z a, Z b)	// acc += a b
{
	acc[0] += (a.x * b.x - a.y * b.y);
	acc[1] += (a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ bool bad(double v) { return v != v; }

// ======================================================================================== epilogues
// alpha = rho / <w, A x> (conj-first).  BiCG w = d2 (clcg.cpp:170-172); CGS/BICGSTAB/TFQMR w = r0bar (:463-465, :620-622, :759-762)
struct EpiInnerAlpha {
	static constexpr int NRED = 2;
	static constexpr bool ACTIVE = true;
	const ZV* w;
	__device__ void begin(const DevState*) {}
	__device__ void row(int i, Z yi, Z, double* acc) const { acc_inner(acc, w[i], yi); }
	__device__ void finish(DevState* st, const double* tot) const { sc_stz(st, SC_ALPHA, zdiv(sc_ldz(st, SC_RHO), zmk(tot[0], tot[1]))); }
};

// alpha = rho / (d . A d) (unconjugated).  BICG_SYM clcg.cpp:319-321, PCG clcg_cuda.cu:514-516
struct EpiDotuAlpha {
	static constexpr int NRED = 2;
	static constexpr bool ACTIVE = true;
	__device__ void begin(const DevState*) {}
	__device__ void row(int, Z yi, Z xi, double* acc) const { acc_dotu(acc, xi, yi); }
	__device__ void finish(DevState* st, const double* tot) const { sc_stz(st, SC_ALPHA, zdiv(sc_ldz(st, SC_RHO), zmk(tot[0], tot[1]))); }
};

// omega = <As, s> / <As, As> (clcg.cpp:630-633)
struct EpiCOmega {
	static constexpr int NRED = 3;
	static constexpr bool ACTIVE = true;
	__device__ void begin(const DevState*) {}
	__device__ void row(int, Z yi, Z xi, double* acc) const { acc_inner(acc, yi, xi); acc[2] += (yi.x * yi.x + yi.y * yi.y); }
	__device__ void finish(DevState* st, const double* tot) const { sc_stz(st, SC_OMEGA, zdiv(zmk(tot[0], tot[1]), zmk(tot[2], 0.0))); }
};

// ======================================================================================== BiCG (clcg.cpp:77-226)
struct OpCbInit : OpBase {	// d1 = r1 = B - Ax, d2 = r2 = conj(r1); <r2,r1>, m.m, r.r (clcg.cpp:102-121) + first head
	static constexpr int NRED = 4, W = 1;
	const ZV* m; const ZV* Ax; const ZV* B; ZV* r1; ZV* r2; ZV* d1; ZV* d2;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z r = zsub(B[i], Ax[i]), rc = zconj(r), mi = m[i];
		r1[i] = r; d1[i] = r; r2[i] = rc; d2[i] = rc;
		acc_inner(acc, rc, r);
		acc[2] += znorm2(mi); acc[3] += znorm2(r);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		sc_stz(st, SC_RHO, zmk(tot[0], tot[1]));
		st->sc[SC_MMOD] = tot[2]; st->sc[SC_RMOD] = tot[3];
		first_head_cplx(st, tot[3], tot[2]);
	}
};

struct OpCbUpdate1 : OpBase {	// m += a d1, r1 -= a A d1; m.m, r.r (clcg.cpp:174-186)
	static constexpr int NRED = 2, W = 1;
	ZV* m; const ZV* d1; ZV* r1; const ZV* Ax; Z ak;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z mi = zadd(m[i], zmul(ak, d1[i])), ri = zsub(r1[i], zmul(ak, Ax[i]));
		m[i] = mi; r1[i] = ri;
		acc[0] += znorm2(mi); acc[1] += znorm2(ri);
	}
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_MMOD] = tot[0]; st->sc[SC_RMOD] = tot[1]; }
};

struct OpCbUpdate2 : OpBase {	// r2 -= conj(a) A^H d2; NaN; <r2,r1>; beta (clcg.cpp:190-206) + head
	static constexpr int NRED = 2, W = 1;
	ZV* r2; const ZV* Ax; const ZV* r1; Z akc;
	__device__ void begin(const DevState* st) { akc = zconj(sc_ldz(st, SC_ALPHA)); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z r = zsub(r2[i], zmul(akc, Ax[i]));
		r2[i] = r;
		acc_inner(acc, r, r1[i]);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (bad(st->sc[SC_MMOD])) { st->ret = RC_C_NAN; st->done = 1; return; }
		Z nxt = zmk(tot[0], tot[1]);
		sc_stz(st, SC_BETA, zdiv(nxt, sc_ldz(st, SC_RHO)));
		sc_stz(st, SC_RHO, nxt);
		loop_head_cplx(st, st->sc[SC_RMOD], st->sc[SC_MMOD]);
	}
};

struct OpCbDir : OpBase {	// d1 = r1 + b d1, d2 = r2 + conj(b) d2 (clcg.cpp:208-213)
	static constexpr int NRED = 0, W = 1;
	const ZV* r1; const ZV* r2; ZV* d1; ZV* d2; Z bk;
	__device__ void begin(const DevState* st) { bk = sc_ldz(st, SC_BETA); }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		d1[i] = zadd(r1[i], zmul(bk, d1[i]));
		d2[i] = zadd(r2[i], zmul(zconj(bk), d2[i]));
	}
};

static int run_cbicg(Engine& E, const Operator<ZV>& A, ZV* m, const ZV* B, size_t n, size_t next)
{
	ZV* r1 = E.alloc<ZV>(next); ZV* r2 = E.alloc<ZV>(next); ZV* d1 = E.alloc<ZV>(next); ZV* d2 = E.alloc<ZV>(next); ZV* Ax = E.alloc<ZV>(next);
	E.spmv(A, m, Ax, EpiNone<ZV>{});
	E.vec_push(OpCbInit{{}, m, Ax, B, r1, r2, d1, d2}, n, d1);
	std::function<void(int)> batch;
	if (E.small_system(A) && A.h->lpr == A.h->t_lpr)
		batch = [&](int k) { E.fused(k, 2, E.ph_spmv(A, d1, Ax, EpiInnerAlpha{d2}), E.ph_vec(OpCbUpdate1{{}, m, d1, r1, Ax, zc()}, n), E.ph_spmv_h(A, d2, Ax, EpiNone<ZV>{}),
			E.ph_vec(OpCbUpdate2{{}, r2, Ax, r1, zc()}, n), E.ph_vec(OpCbDir{{}, r1, r2, d1, d2, zc()}, n)); };
	return E.run([&]() {
		E.spmv(A, d1, Ax, EpiInnerAlpha{d2});
		E.vec(OpCbUpdate1{{}, m, d1, r1, Ax, zc()}, n);
		E.spmv(A, d2, Ax, EpiNone<ZV>{}, 2);	// A^H d2 (MatTranspose, Conjugate — clcg.cpp:188)
		E.vec2_push(OpCbUpdate2{{}, r2, Ax, r1, zc()}, OpCbDir{{}, r1, r2, d1, d2, zc()}, n, d1);
		return false;
	}, batch);
}

// ====================================================== BICG_SYM (clcg.cpp:228-364) and PCG (clcg_cuda.cu:403-559)
// MODE 0: symmetric BiCG (rho = r.r, d = r);  1: PCG with fused Jacobi z = r/diag (rho = r.z, d = z);
//      2: PCG, user preconditioner: this kernel only does the m/r part, OpCsRho finishes after the callback.
template <int MODE>
struct OpCsInit : OpBase {
	static constexpr int NRED = 4, W = 1;
	const ZV* m; const ZV* Ax; const ZV* B; const ZV* diag; ZV* r; ZV* z; ZV* d;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z ri = zsub(B[i], Ax[i]), mi = m[i];
		r[i] = ri;
		acc[2] += znorm2(mi); acc[3] += znorm2(ri);
		if (MODE == 0) { d[i] = ri; acc_dotu(acc, ri, ri); }
		if (MODE == 1) { Z zi = zdiv(ri, diag[i]); z[i] = zi; d[i] = zi; acc_dotu(acc, ri, zi); }
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		st->sc[SC_MMOD] = tot[2]; st->sc[SC_RMOD] = tot[3];
		if (MODE == 2) return;
		sc_stz(st, SC_RHO, zmk(tot[0], tot[1]));
		first_head_cplx(st, tot[3], tot[2]);
	}
};

struct OpCsInitZ : OpBase {	// MODE 2: d = z, rho = r.z, then the first head
	static constexpr int NRED = 2, W = 1;
	const ZV* r; const ZV* z; ZV* d;
	template <int V> __device__ void elem(size_t i, double* acc) const { Z zi = z[i]; d[i] = zi; acc_dotu(acc, r[i], zi); }
	__device__ void finish(DevState* st, const double* tot) const
	{
		sc_stz(st, SC_RHO, zmk(tot[0], tot[1]));
		first_head_cplx(st, st->sc[SC_RMOD], st->sc[SC_MMOD]);
	}
};

__device__ __forceinline__ void cs_tail(DevState* st, Z rho2)
{
	// NaN exit: clcg.cpp:337-343 for BICG_SYM.  The reference PCG loops have no NaN test and would spin forever
	// on a NaN iterate when max_iterations = 0; we return CLCG_NAN_VALUE there too (documented deviation).
	if (bad(st->sc[SC_MMOD])) { st->ret = RC_C_NAN; st->done = 1; return; }
	sc_stz(st, SC_BETA, zdiv(rho2, sc_ldz(st, SC_RHO)));
	sc_stz(st, SC_RHO, rho2);
	loop_head_cplx(st, st->sc[SC_RMOD], st->sc[SC_MMOD]);
}

template <int MODE>
struct OpCsUpdate : OpBase {	// m += a d, r -= a Ad [, z = r/diag]; m.m, r.r, rho' (clcg.cpp:323-347 / clcg_cuda.cu:518-539)
	static constexpr int NRED = 4, W = 1;
	ZV* m; const ZV* d; ZV* r; const ZV* Ax; const ZV* diag; ZV* z; Z ak;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z mi = zadd(m[i], zmul(ak, d[i])), ri = zsub(r[i], zmul(ak, Ax[i]));
		m[i] = mi; r[i] = ri;
		acc[2] += znorm2(mi); acc[3] += znorm2(ri);
		if (MODE == 0) acc_dotu(acc, ri, ri);
		if (MODE == 1) { Z zi = zdiv(ri, diag[i]); z[i] = zi; acc_dotu(acc, ri, zi); }
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		st->sc[SC_MMOD] = tot[2]; st->sc[SC_RMOD] = tot[3];
		if (MODE == 2) return;
		cs_tail(st, zmk(tot[0], tot[1]));
	}
};

struct OpCsRho : OpBase {	// MODE 2: rho' = r.z after the user's preconditioner
	static constexpr int NRED = 2, W = 1;
	const ZV* r; const ZV* z;
	template <int V> __device__ void elem(size_t i, double* acc) const { acc_dotu(acc, r[i], z[i]); }
	__device__ void finish(DevState* st, const double* tot) const { cs_tail(st, zmk(tot[0], tot[1])); }
};

struct OpCsDir : OpBase {	// d = s + b d with s = r (BICG_SYM, clcg.cpp:349-353) or s = z (PCG, clcg_cuda.cu:536-537)
	static constexpr int NRED = 0, W = 1;
	const ZV* s; ZV* d; Z bk;
	__device__ void begin(const DevState* st) { bk = sc_ldz(st, SC_BETA); }
	template <int V> __device__ void elem(size_t i, double*) const { d[i] = zadd(s[i], zmul(bk, d[i])); }
};

static int run_csym(Engine& E, const Operator<ZV>& A, ZV* m, const ZV* B, size_t n, size_t next, bool pcg)
{
	ZV* r = E.alloc<ZV>(next); ZV* d = E.alloc<ZV>(next); ZV* Ax = E.alloc<ZV>(next);
	ZV* z = pcg ? E.alloc<ZV>(next) : nullptr;
	const int mode = !pcg ? 0 : (A.diag ? 1 : 2);
	E.spmv(A, m, Ax, EpiNone<ZV>{});
	if (mode == 0) E.vec_push(OpCsInit<0>{{}, m, Ax, B, nullptr, r, z, d}, n, d);
	else if (mode == 1) E.vec_push(OpCsInit<1>{{}, m, Ax, B, A.diag, r, z, d}, n, d);
	else
	{
		E.vec(OpCsInit<2>{{}, m, Ax, B, nullptr, r, z, d}, n);
		E.precondition(A, r, z);
		E.vec_push(OpCsInitZ{{}, r, z, d}, n, d);
	}
	std::function<void(int)> batch;
	if (E.small_system(A) && mode == 0)
		batch = [&](int k) { E.fused(k, 1, E.ph_spmv(A, d, Ax, EpiDotuAlpha{}), E.ph_vec(OpCsUpdate<0>{{}, m, d, r, Ax, nullptr, z, zc()}, n), E.ph_vec(OpCsDir{{}, r, d, zc()}, n)); };
	if (E.small_system(A) && mode == 1)
		batch = [&](int k) { E.fused(k, 1, E.ph_spmv(A, d, Ax, EpiDotuAlpha{}), E.ph_vec(OpCsUpdate<1>{{}, m, d, r, Ax, A.diag, z, zc()}, n), E.ph_vec(OpCsDir{{}, z, d, zc()}, n)); };
	return E.run([&]() {
		E.spmv(A, d, Ax, EpiDotuAlpha{});
		if (mode == 0) E.vec2_push(OpCsUpdate<0>{{}, m, d, r, Ax, nullptr, z, zc()}, OpCsDir{{}, r, d, zc()}, n, d);
		else if (mode == 1) E.vec2_push(OpCsUpdate<1>{{}, m, d, r, Ax, A.diag, z, zc()}, OpCsDir{{}, z, d, zc()}, n, d);
		else
		{
			E.vec(OpCsUpdate<2>{{}, m, d, r, Ax, nullptr, z, zc()}, n);
			E.precondition(A, r, z);
			E.vec2_push(OpCsRho{{}, r, z}, OpCsDir{{}, z, d, zc()}, n, d);
		}
		return false;
	}, batch);
}

// ======================================================================================== shadow residual
// r0bar exactly as the reference draws it (lcg_complex.cpp:118-127 with l = 1+0i, h = 2+0i): libc srand/rand on
// the host, the imaginary draw consumes a rand() too.  `skip` = global index of the first local element.
static void draw_shadow(std::vector<ZV>& host, size_t n, long seed, size_t skip)
{
	srand((unsigned)(seed ? seed : (long)time(nullptr)));
	for (size_t i = 0; i < skip; i++) { (void)rand(); (void)rand(); }
	host.resize(n);
	for (size_t i = 0; i < n; i++)
	{
		double re = (2.0 - 1.0) * rand() * 1.0 / RAND_MAX + 1.0;
		double im = (0.0 - 0.0) * rand() * 1.0 / RAND_MAX + 0.0;
		host[i] = make_double2(re, im);
	}
}

struct OpCResInit : OpBase {	// p = [u =] r = B - Ax [, d = 0]; m.m, r.r (clcg.cpp:392-396, 549-553, 709-719)
	static constexpr int NRED = 2, W = 1;
	const ZV* m; const ZV* Ax; const ZV* B; ZV* r; ZV* p; ZV* u; ZV* dz;
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z ri = zsub(B[i], Ax[i]);
		r[i] = ri; p[i] = ri;
		if (u) u[i] = ri;
		if (dz) dz[i] = zmk(0.0, 0.0);
		acc[0] += znorm2(m[i]); acc[1] += znorm2(ri);
	}
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_MMOD] = tot[0]; st->sc[SC_RMOD] = tot[1]; }
};

// rho = <r0bar, r>; when |rho| >= 1e-8 go on to the "already optimised" test and (HEAD) the first loop head.
// flag = 1 asks the host to redraw r0bar (clcg.cpp:399-403).
template <bool HEAD>
struct OpCRho : OpBase {
	static constexpr int NRED = 2, W = 1;
	const ZV* rb; const ZV* r;
	template <int V> __device__ void elem(size_t i, double* acc) const { acc_inner(acc, rb[i], r[i]); }
	__device__ void finish(DevState* st, const double* tot) const
	{
		Z rho = zmk(tot[0], tot[1]);
		if (sqrt(znorm2(rho)) < 1e-8) { st->flag = 1; return; }
		st->flag = 0;
		sc_stz(st, SC_RHO, rho);
		if (HEAD) first_head_cplx(st, st->sc[SC_RMOD], st->sc[SC_MMOD]);
		else
		{	// TFQMR: only the "already optimised" test here; its loop heads sit inside the half steps
			double r;
			bool hit = false;
			if (st->abs_diff && (r = cplx_res_abs(st, st->sc[SC_RMOD])) <= st->eps) hit = true;
			else if ((r = cplx_res_rel(st, st->sc[SC_RMOD], st->sc[SC_MMOD])) <= st->eps) hit = true;
			if (hit) { st->residual = r; st->k_report = 0; st->checks++; st->ret = RC_ALREADY; st->done = 1; }
		}
	}
};

// Partitioned solves with the reference's clock seed (shadow_seed = 0): every rank must seed rand() identically, so the ranks
// agree on the mean of their clocks through the same cross-rank sum the dot products use (bitwise identical on all ranks).
struct OpSeedAgree : OpBase {
	static constexpr int NRED = 1, W = 1;
	double t;
	template <int V> __device__ void elem(size_t i, double* acc) const { if (i == 0) acc[0] += t; }
	__device__ void finish(DevState* st, const double* tot) const { st->sc[SC_TMP5] = tot[0]; }
};

template <bool HEAD>
static bool init_shadow(Engine& E, ZV* rb, const ZV* r, size_t n, size_t skip)
{
	std::vector<ZV> host;
	long seed = settings().shadow_seed;
	if (seed == 0 && E.multi())
	{
		E.vec(OpSeedAgree{{}, (double)time(nullptr)}, 1);
		E.read_state();
		seed = (long)(E.h_st->sc[SC_TMP5] / (double)E.comm->size());
		if (seed == 0) seed = 1;
	}
	for (int attempt = 0; attempt < 64; attempt++)
	{
		draw_shadow(host, n, seed, skip);
		LCG_CUDA_CHECK(cudaMemcpyAsync(rb, host.data(), n * sizeof(ZV), cudaMemcpyHostToDevice, E.stream));
		E.vec(OpCRho<HEAD>{{}, rb, r}, n);
		E.read_state();	// also makes the pageable `host` buffer safe to reuse
		if (!E.h_st->flag) return true;
		seed = (seed ? seed : (long)time(nullptr)) + 1;	// the reference re-seeds from the clock until it ticks
	}
	set_error_msg("could not draw a shadow residual with |<r0bar,r>| >= 1e-8");
	return false;
}

// ======================================================================================== CGS (clcg.cpp:366-522)
struct OpCQW : OpBase {	// q = u - a Ap, w = u + q (clcg.cpp:467-472, 764-769)
	static constexpr int NRED = 0, W = 1;
	const ZV* u; const ZV* Ax; ZV* q; ZV* w; Z ak;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		Z ui = u[i], qi = zsub(ui, zmul(ak, Ax[i]));
		q[i] = qi; w[i] = zadd(ui, qi);
	}
};

struct OpCCgsUpdate : OpBase {	// m += a w, r -= a Aw; m.m, r.r, <r0bar,r>; beta (clcg.cpp:476-500) + head
	static constexpr int NRED = 4, W = 1;
	ZV* m; const ZV* w; ZV* r; const ZV* Ax; const ZV* rb; Z ak;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z mi = zadd(m[i], zmul(ak, w[i])), ri = zsub(r[i], zmul(ak, Ax[i]));
		m[i] = mi; r[i] = ri;
		acc_inner(acc, rb[i], ri);
		acc[2] += znorm2(mi); acc[3] += znorm2(ri);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (bad(tot[2])) { st->ret = RC_C_NAN; st->done = 1; return; }
		st->sc[SC_MMOD] = tot[2]; st->sc[SC_RMOD] = tot[3];
		Z rho2 = zmk(tot[0], tot[1]);
		sc_stz(st, SC_BETA, zdiv(rho2, sc_ldz(st, SC_RHO)));
		sc_stz(st, SC_RHO, rho2);
		loop_head_cplx(st, tot[3], tot[2]);
	}
};

struct OpCCgsDir : OpBase {	// u = r + b q, p = u + b (q + b p) (clcg.cpp:502-507, 860-865)
	static constexpr int NRED = 0, W = 1;
	const ZV* r; const ZV* q; ZV* u; ZV* p; Z bk;
	__device__ void begin(const DevState* st) { bk = sc_ldz(st, SC_BETA); }
	template <int V> __device__ void elem(size_t i, double*) const
	{
		Z qi = q[i], ui = zadd(r[i], zmul(bk, qi));
		u[i] = ui;
		p[i] = zadd(ui, zmul(bk, zadd(qi, zmul(bk, p[i]))));
	}
};

static int run_ccgs(Engine& E, const Operator<ZV>& A, ZV* m, const ZV* B, size_t n, size_t next, size_t skip)
{
	ZV* r = E.alloc<ZV>(next); ZV* rb = E.alloc<ZV>(next); ZV* p = E.alloc<ZV>(next); ZV* Ax = E.alloc<ZV>(next);
	ZV* u = E.alloc<ZV>(next); ZV* q = E.alloc<ZV>(next); ZV* w = E.alloc<ZV>(next);
	E.spmv(A, m, Ax, EpiNone<ZV>{});
	E.vec_push(OpCResInit{{}, m, Ax, B, r, p, u, nullptr}, n, p);
	if (!init_shadow<true>(E, rb, r, n, skip)) return RC_UNKNOWN;
	std::function<void(int)> batch;
	if (E.small_system(A))
		batch = [&](int k) { E.fused(k, 2, E.ph_spmv(A, p, Ax, EpiInnerAlpha{rb}), E.ph_vec(OpCQW{{}, u, Ax, q, w, zc()}, n), E.ph_spmv(A, w, Ax, EpiNone<ZV>{}),
			E.ph_vec(OpCCgsUpdate{{}, m, w, r, Ax, rb, zc()}, n), E.ph_vec(OpCCgsDir{{}, r, q, u, p, zc()}, n)); };
	return E.run([&]() {
		E.spmv(A, p, Ax, EpiInnerAlpha{rb});
		E.vec_push(OpCQW{{}, u, Ax, q, w, zc()}, n, w);
		E.spmv(A, w, Ax, EpiNone<ZV>{});
		E.vec2_push(OpCCgsUpdate{{}, m, w, r, Ax, rb, zc()}, OpCCgsDir{{}, r, q, u, p, zc()}, n, p);
		return false;
	}, batch);
}

// ======================================================================================== BICGSTAB (clcg.cpp:524-679)
struct OpCBsS : OpBase {	// s = r - a Ap (clcg.cpp:624-628)
	static constexpr int NRED = 0, W = 1;
	const ZV* r; const ZV* Ap; ZV* s; Z ak;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); }
	template <int V> __device__ void elem(size_t i, double*) const { s[i] = zsub(r[i], zmul(ak, Ap[i])); }
};

struct OpCBsUpdate : OpBase {	// m += a p + w s, r = s - w As; m.m, r.r, <r0bar,r>; beta (clcg.cpp:635-659) + head
	static constexpr int NRED = 4, W = 1;
	ZV* m; const ZV* p; const ZV* s; const ZV* As; ZV* r; const ZV* rb; Z ak, wk;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); wk = sc_ldz(st, SC_OMEGA); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z si = s[i];
		Z mi = zadd(zadd(m[i], zmul(ak, p[i])), zmul(wk, si));
		Z ri = zsub(si, zmul(wk, As[i]));
		m[i] = mi; r[i] = ri;
		acc_inner(acc, rb[i], ri);
		acc[2] += znorm2(mi); acc[3] += znorm2(ri);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (bad(tot[2])) { st->ret = RC_C_NAN; st->done = 1; return; }
		st->sc[SC_MMOD] = tot[2]; st->sc[SC_RMOD] = tot[3];
		Z rho2 = zmk(tot[0], tot[1]);
		// betak = rhok2*ak/(rhok*omega)
		sc_stz(st, SC_BETA, zdiv(zmul(rho2, sc_ldz(st, SC_ALPHA)), zmul(sc_ldz(st, SC_RHO), sc_ldz(st, SC_OMEGA))));
		sc_stz(st, SC_RHO, rho2);
		loop_head_cplx(st, tot[3], tot[2]);
	}
};

struct OpCBsDir : OpBase {	// p = r + b (p - w Ap) (clcg.cpp:661-665)
	static constexpr int NRED = 0, W = 1;
	const ZV* r; ZV* p; const ZV* Ap; Z bk, wk;
	__device__ void begin(const DevState* st) { bk = sc_ldz(st, SC_BETA); wk = sc_ldz(st, SC_OMEGA); }
	template <int V> __device__ void elem(size_t i, double*) const { p[i] = zadd(r[i], zmul(bk, zsub(p[i], zmul(wk, Ap[i])))); }
};

static int run_cbicgstab(Engine& E, const Operator<ZV>& A, ZV* m, const ZV* B, size_t n, size_t next, size_t skip)
{
	ZV* r = E.alloc<ZV>(next); ZV* rb = E.alloc<ZV>(next); ZV* p = E.alloc<ZV>(next); ZV* s = E.alloc<ZV>(next);
	ZV* Ap = E.alloc<ZV>(next); ZV* As = E.alloc<ZV>(next);
	E.spmv(A, m, Ap, EpiNone<ZV>{});
	E.vec_push(OpCResInit{{}, m, Ap, B, r, p, nullptr, nullptr}, n, p);
	if (!init_shadow<true>(E, rb, r, n, skip)) return RC_UNKNOWN;
	std::function<void(int)> batch;
	if (E.small_system(A))
		batch = [&](int k) { E.fused(k, 2, E.ph_spmv(A, p, Ap, EpiInnerAlpha{rb}), E.ph_vec(OpCBsS{{}, r, Ap, s, zc()}, n), E.ph_spmv(A, s, As, EpiCOmega{}),
			E.ph_vec(OpCBsUpdate{{}, m, p, s, As, r, rb, zc(), zc()}, n), E.ph_vec(OpCBsDir{{}, r, p, Ap, zc(), zc()}, n)); };
	return E.run([&]() {
		E.spmv(A, p, Ap, EpiInnerAlpha{rb});
		E.vec_push(OpCBsS{{}, r, Ap, s, zc()}, n, s);
		E.spmv(A, s, As, EpiCOmega{});
		E.vec2_push(OpCBsUpdate{{}, m, p, s, As, r, rb, zc(), zc()}, OpCBsDir{{}, r, p, Ap, zc(), zc()}, n, p);
		return false;
	}, batch);
}

// ======================================================================================== TFQMR (clcg.cpp:681-882)
// Scalars of a half step (clcg.cpp:808-833); j = 1 or 2.  SC_RKM = |<r,r>| of the previous outer iteration,
// SC_RKM2 = |<r,r>| after this outer iteration's residual update.
__device__ __forceinline__ void tfqmr_half_scalars(DevState* st, int j)
{
	const double theta0 = st->sc[SC_THETA];
	Z eta = sc_ldz(st, SC_ETA), alpha = sc_ldz(st, SC_ALPHA);
	Z sign = zscale(theta0 * theta0, zdiv(eta, alpha));
	sc_stz(st, SC_TMP0, sign);
	const double omega = (j == 1) ? sqrt(st->sc[SC_RKM] * st->sc[SC_RKM2]) : st->sc[SC_RKM2];
	const double theta = omega / st->sc[SC_TAO];
	st->sc[SC_THETA] = theta;
	st->sc[SC_TAO] = omega / sqrt(1.0 + theta * theta);
	sc_stz(st, SC_ETA, zscale(1.0 / (1.0 + theta * theta), alpha));
	st->half = j;
}

struct OpCTfR : OpBase {	// r -= a A(u+q); <r,r>, <r0bar,r> (clcg.cpp:773-779, 856) + head of half step 1
	static constexpr int NRED = 3, W = 1;
	ZV* r; const ZV* Ax; const ZV* rb; Z ak;
	__device__ void begin(const DevState* st) { ak = sc_ldz(st, SC_ALPHA); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z ri = zsub(r[i], zmul(ak, Ax[i]));
		r[i] = ri;
		acc_inner(acc, rb[i], ri);
		acc[2] += znorm2(ri);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		st->sc[SC_RKM2] = tot[2];
		sc_stz(st, SC_TMP2, zmk(tot[0], tot[1]));	// rho2, consumed after the second half step
		loop_head_cplx(st, st->sc[SC_RKM], st->sc[SC_MMOD]);	// rk_square is still the previous outer iteration's
		if (!st->done) tfqmr_half_scalars(st, 1);
	}
};

template <int J>
struct OpCTfHalf : OpBase {	// d = (u|q) + sign d, m += eta d; m.m, NaN (clcg.cpp:810-851) [+ head of half step 2]
	static constexpr int NRED = 1, W = 1;
	const ZV* uq; ZV* d; ZV* m; Z sign, eta;
	__device__ void begin(const DevState* st) { sign = sc_ldz(st, SC_TMP0); eta = sc_ldz(st, SC_ETA); }
	template <int V> __device__ void elem(size_t i, double* acc) const
	{
		Z di = zadd(uq[i], zmul(sign, d[i]));
		d[i] = di;
		Z mi = zadd(m[i], zmul(eta, di));
		m[i] = mi;
		acc[0] += znorm2(mi);
	}
	__device__ void finish(DevState* st, const double* tot) const
	{
		if (bad(tot[0])) { st->ret = RC_C_NAN; st->done = 1; return; }
		st->sc[SC_MMOD] = tot[0];
		if (J == 1)
		{
			loop_head_cplx(st, st->sc[SC_RKM], tot[0]);
			if (!st->done) tfqmr_half_scalars(st, 2);
		}
		else
		{	// end of the outer iteration: rk_mod = rk_mod2, beta = rho2/rho (clcg.cpp:853-858)
			st->sc[SC_RKM] = st->sc[SC_RKM2];
			Z rho2 = sc_ldz(st, SC_TMP2);
			sc_stz(st, SC_BETA, zdiv(rho2, sc_ldz(st, SC_RHO)));
			sc_stz(st, SC_RHO, rho2);
		}
	}
};

static int run_ctfqmr(Engine& E, const Operator<ZV>& A, ZV* m, const ZV* B, size_t n, size_t next, size_t skip)
{
	ZV* p = E.alloc<ZV>(next); ZV* u = E.alloc<ZV>(next); ZV* v = E.alloc<ZV>(next); ZV* d = E.alloc<ZV>(next);
	ZV* rb = E.alloc<ZV>(next); ZV* r = E.alloc<ZV>(next); ZV* Ax = E.alloc<ZV>(next); ZV* q = E.alloc<ZV>(next); ZV* uq = E.alloc<ZV>(next);
	E.spmv(A, m, Ax, EpiNone<ZV>{});
	E.vec_push(OpCResInit{{}, m, Ax, B, r, p, u, d}, n, p);
	if (!init_shadow<false>(E, rb, r, n, skip)) return RC_UNKNOWN;
	// theta = 0, omega = tao = |<r,r>| = r.r, eta = 0 — written straight into the state block
	{
		double init[2] = {0.0, 0.0};
		char* base = reinterpret_cast<char*>(E.d_st) + offsetof(DevState, sc);
		LCG_CUDA_CHECK(cudaMemcpyAsync(base + sizeof(double) * SC_THETA, init, sizeof(double), cudaMemcpyHostToDevice, E.stream));
		LCG_CUDA_CHECK(cudaMemcpyAsync(base + sizeof(double) * SC_ETA, init, 2 * sizeof(double), cudaMemcpyHostToDevice, E.stream));
		LCG_CUDA_CHECK(cudaMemcpyAsync(base + sizeof(double) * SC_TAO, base + sizeof(double) * SC_RMOD, sizeof(double), cudaMemcpyDeviceToDevice, E.stream));
		LCG_CUDA_CHECK(cudaMemcpyAsync(base + sizeof(double) * SC_RKM, base + sizeof(double) * SC_RMOD, sizeof(double), cudaMemcpyDeviceToDevice, E.stream));
		LCG_CUDA_CHECK(cudaStreamSynchronize(E.stream));	// `init` is a stack buffer
	}
	return E.run([&]() {
		E.spmv(A, p, v, EpiInnerAlpha{rb});
		E.vec_push(OpCQW{{}, u, v, q, uq, zc()}, n, uq);
		E.spmv(A, uq, Ax, EpiNone<ZV>{});
		E.vec(OpCTfR{{}, r, Ax, rb, zc()}, n);
		if (E.sync_point()) return true;
		E.vec(OpCTfHalf<1>{{}, u, d, m, zc(), zc()}, n);
		if (E.sync_point()) return true;
		E.vec(OpCTfHalf<2>{{}, q, d, m, zc(), zc()}, n);
		E.vec_push(OpCCgsDir{{}, r, q, u, p, zc()}, n, p);
		return false;
	});
}

// ======================================================================================== dispatch
static int dispatch(Engine& E, const Operator<ZV>& A, int solver_id, ZV* m, const ZV* B, const lcgb200_cpara& para, size_t n, size_t next)
{
	const size_t skip = A.h ? (size_t)A.h->row_offset : 0;	// partitioned: global index of the first local row (lcgb200_csr_set_row_offset)
	switch (solver_id)
	{
		case LCGB200_CBICG: return run_cbicg(E, A, m, B, n, next);
		case LCGB200_CBICG_SYM: return run_csym(E, A, m, B, n, next, false);
		case LCGB200_CPCG: return run_csym(E, A, m, B, n, next, true);
		case LCGB200_CBICGSTAB: return run_cbicgstab(E, A, m, B, n, next, skip);
		case LCGB200_CTFQMR: return run_ctfqmr(E, A, m, B, n, next, skip);
		case LCGB200_CCGS: default: return run_ccgs(E, A, m, B, n, next, skip);
	}
}

}  // namespace LCG_CPLX_NS

#ifdef LCG_CPLX_FLOAT
int solve_complexf(Engine& E, const Operator<ZF>& A, int solver_id, ZF* m, const ZF* B, const lcgb200_cpara& para, size_t n, size_t next)
{
	return cf32::dispatch(E, A, solver_id, m, B, para, n, next);
}
#elif defined(LCG_REFORDER)
int solve_complex_x(Engine& E, const Operator<double2>& A, int solver_id, double2* m, const double2* B, const lcgb200_cpara& para, size_t n, size_t next)
{
	return cx64::dispatch(E, A, solver_id, m, B, para, n, next);
}
#else
int solve_complex(Engine& E, const Operator<double2>& A, int solver_id, double2* m, const double2* B, const lcgb200_cpara& para, size_t n, size_t next)
{
	return cf64::dispatch(E, A, solver_id, m, B, para, n, next);
}

int complex_vector_count(int solver_id)
{
	switch (solver_id)
	{
		case LCGB200_CBICG: return 5;
		case LCGB200_CBICG_SYM: return 3;
		case LCGB200_CPCG: return 4;
		case LCGB200_CBICGSTAB: return 6;
		case LCGB200_CTFQMR: return 9;
		default: return 7;
	}
}

#endif

}  // namespace lcgb200
