// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// Thin extern "C" driver around the UNMODIFIED reference CPU/OpenMP library (compiled from
// /root/reference/src/lib/{algebra,util,lcg,lcg_complex,clcg}.cpp by oracle/Makefile into
// oracle/_ref/liblcg_ref.so).  It supplies what the reference's own samples supply by hand:
// an OpenMP CSR `Ax` callback, a Jacobi `M^-1 x` callback and a progress callback that records
// the (k, residual) history — the same roles as cudaAx / cudaMx / cudaProgress in
// /root/reference/src/sample/sample8.cu:96-128 and MxProduct in sample10.cu:100-121.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include <complex>
#include <dlfcn.h>
#include <omp.h>

#include "lcg.h"   // resolved with -I/root/reference/src/lib at build time (never copied)
#include "clcg.h"

// ---------------------------------------------------------------------------------------------
// Deterministic seed for clcg_vecrnd(): the reference seeds rand() with time(0) on every call
// (/root/reference/src/lib/lcg_complex.cpp:118-127).  This .so is linked -Bsymbolic, so the
// reference objects inside it bind to this definition of time(); nothing outside is affected.
static long g_fixed_time = 0;
extern "C" time_t time(time_t* t)
{
	time_t v;
	if (g_fixed_time != 0) v = (time_t)g_fixed_time;
	else
	{
		typedef time_t (*time_fn)(time_t*);
		static time_fn real_time = (time_fn)dlsym(RTLD_NEXT, "time");
		v = real_time ? real_time(nullptr) : (time_t)0;
	}
	if (t) *t = v;
	return v;
}

extern "C" void lcgref_set_time(long fixed) { g_fixed_time = fixed; }
extern "C" int lcgref_num_threads() { return omp_get_max_threads(); }
extern "C" void lcgref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// ---------------------------------------------------------------------------------------------
struct History
{
	int last_k = -1;
	double last_res = 0.0;
	int calls = 0;
	double* buf = nullptr;
	int cap = 0;
	int stop_at = -1;  // Pfp returns 1 when k == stop_at (exercises LCG_STOP)
	void push(int k, double r)
	{
		last_k = k; last_res = r;
		if (buf && calls < cap) buf[calls] = r;
		calls++;
	}
};

struct RealSys
{
	int n;
	const int* rp; const int* ci; const double* v;
	const double* diag;
	History h;
};

static void real_ax(void* inst, const lcg_float* x, lcg_float* y, const int n)
{
	RealSys* s = (RealSys*)inst;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++)
	{
		double acc = 0.0;
		for (int k = s->rp[i]; k < s->rp[i + 1]; k++) acc += s->v[k] * x[s->ci[k]];
		y[i] = acc;
	}
}

static void real_mx(void* inst, const lcg_float* x, lcg_float* y, const int n)
{
	RealSys* s = (RealSys*)inst;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) y[i] = x[i] / s->diag[i];
}

static int real_pf(void* inst, const lcg_float* m, const lcg_float conv, const lcg_para* p, const int n, const int k)
{
	RealSys* s = (RealSys*)inst;
	s->h.push(k, conv);
	return (s->h.stop_at >= 0 && k == s->h.stop_at) ? 1 : 0;
}

// out[0]=last k seen by Pfp, out[1]=number of Pfp calls; dout[0]=last residual, dout[1]=seconds in solver
extern "C" int lcgref_solve(int solver_id, int n, const int* rp, const int* ci, const double* val,
	double* m, const double* B, const double* low, const double* hig, const double* diag,
	const void* para, int use_progress, int stop_at, double* hist, int hist_cap, int* out, double* dout)
{
	RealSys s; s.n = n; s.rp = rp; s.ci = ci; s.v = val; s.diag = diag;
	s.h.buf = hist; s.h.cap = hist_cap; s.h.stop_at = stop_at;
	const lcg_para* p = (const lcg_para*)para;
	lcg_progress_ptr pf = use_progress ? real_pf : nullptr;
	int ret;
	double t0 = omp_get_wtime();
	if (solver_id == LCG_PCG) ret = lcg_solver_preconditioned(real_ax, real_mx, pf, m, B, n, p, &s, LCG_PCG);
	else if (solver_id == LCG_PG || solver_id == LCG_SPG)
		ret = lcg_solver_constrained(real_ax, pf, m, B, low, hig, n, p, &s, (lcg_solver_enum)solver_id);
	else ret = lcg_solver(real_ax, pf, m, B, n, p, &s, (lcg_solver_enum)solver_id);
	double t1 = omp_get_wtime();
	if (out) { out[0] = s.h.last_k; out[1] = s.h.calls; }
	if (dout) { dout[0] = s.h.last_res; dout[1] = t1 - t0; }
	return ret;
}

extern "C" void lcgref_spmv(int n, const int* rp, const int* ci, const double* val, const double* x, double* y)
{
	RealSys s; s.n = n; s.rp = rp; s.ci = ci; s.v = val; s.diag = nullptr;
	real_ax(&s, x, y, n);
}

// ---------------------------------------------------------------------------------------------
typedef std::complex<double> cplx;

struct CplxSys
{
	int n;
	const int* rp; const int* ci; const cplx* v;      // A   (CSR)
	std::vector<int> trp, tci; std::vector<cplx> tv;  // A^T (CSR), built once per solve
	History h;
};

static void build_transpose(CplxSys* s)
{
	int n = s->n, nnz = s->rp[n];
	s->trp.assign(n + 1, 0); s->tci.resize(nnz); s->tv.resize(nnz);
	for (int k = 0; k < nnz; k++) s->trp[s->ci[k] + 1]++;
	for (int i = 0; i < n; i++) s->trp[i + 1] += s->trp[i];
	std::vector<int> fill(s->trp.begin(), s->trp.end() - 1);
	for (int i = 0; i < n; i++)
		for (int k = s->rp[i]; k < s->rp[i + 1]; k++)
		{
			int c = s->ci[k]; int d = fill[c]++;
			s->tci[d] = i; s->tv[d] = s->v[k];
		}
}

static void cplx_ax(void* inst, const lcg_complex* x, lcg_complex* y, const int n, lcg_matrix_e layout, clcg_complex_e conj)
{
	CplxSys* s = (CplxSys*)inst;
	const int* rp = (layout == MatNormal) ? s->rp : s->trp.data();
	const int* ci = (layout == MatNormal) ? s->ci : s->tci.data();
	const cplx* v = (layout == MatNormal) ? s->v : s->tv.data();
	const bool cj = (conj == Conjugate);
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++)
	{
		cplx acc(0.0, 0.0);
		for (int k = rp[i]; k < rp[i + 1]; k++) acc += (cj ? std::conj(v[k]) : v[k]) * x[ci[k]];
		y[i] = acc;
	}
}

static int cplx_pf(void* inst, const lcg_complex* m, const lcg_float conv, const clcg_para* p, const int n, const int k)
{
	CplxSys* s = (CplxSys*)inst;
	s->h.push(k, conv);
	return (s->h.stop_at >= 0 && k == s->h.stop_at) ? 1 : 0;
}

// complex arrays are interleaved (re, im) doubles == std::complex<double> layout
extern "C" int lcgref_csolve(int solver_id, int n, const int* rp, const int* ci, const double* val,
	double* m, const double* B, const void* para, int use_progress, int stop_at,
	double* hist, int hist_cap, int* out, double* dout)
{
	CplxSys s; s.n = n; s.rp = rp; s.ci = ci; s.v = (const cplx*)val;
	s.h.buf = hist; s.h.cap = hist_cap; s.h.stop_at = stop_at;
	build_transpose(&s);
	clcg_progress_ptr pf = use_progress ? cplx_pf : nullptr;
	double t0 = omp_get_wtime();
	int ret = clcg_solver(cplx_ax, pf, (lcg_complex*)m, (const lcg_complex*)B, n, (const clcg_para*)para, &s,
		(clcg_solver_enum)solver_id);
	double t1 = omp_get_wtime();
	if (out) { out[0] = s.h.last_k; out[1] = s.h.calls; }
	if (dout) { dout[0] = s.h.last_res; dout[1] = t1 - t0; }
	return ret;
}

extern "C" void lcgref_cspmv(int n, const int* rp, const int* ci, const double* val, const double* x, double* y,
	int transpose, int conjugate)
{
	CplxSys s; s.n = n; s.rp = rp; s.ci = ci; s.v = (const cplx*)val;
	if (transpose) build_transpose(&s);
	cplx_ax(&s, (const lcg_complex*)x, (lcg_complex*)y, n, transpose ? MatTranspose : MatNormal,
		conjugate ? Conjugate : NonConjugate);
}

// The random shadow residual exactly as the reference draws it (lcg_complex.cpp:118-127), for tests that
// need to know r0bar: calls the reference's own clcg_vecrnd.
extern "C" void lcgref_vecrnd(double* a, int n)
{
	clcg_vecrnd((lcg_complex*)a, lcg_complex(1.0, 0.0), lcg_complex(2.0, 0.0), n);
}

extern "C" int lcgref_sizeof_para() { return (int)sizeof(lcg_para); }
extern "C" int lcgref_sizeof_cpara() { return (int)sizeof(clcg_para); }

// ---------------------------------------------------------------------------------------------
// IC(0) preconditioning with the reference's own functions (preconditioner.cpp): the factorisation
// lcg_incomplete_Cholesky_half_coo and PCG whose Mx callback is the pair of COO triangular solves
// lcg_solve_lower_triangle_coo / lcg_solve_upper_triangle_coo (sample7.cpp does the complex twin of this by hand).
#include "preconditioner.h"

// COO (row-sorted, base 0) of the full matrix -> lower factor in COO; returns the number of entries of L
extern "C" int lcgref_ic0_half(const int* row, const int* col, const double* val, int n, int nz, int* ic_row, int* ic_col, double* ic_val)
{
	int lnz = 0;
	lcg_incomplete_Cholesky_half_buffsize_coo(row, col, nz, &lnz);
	if (ic_row && ic_col && ic_val) lcg_incomplete_Cholesky_half_coo(row, col, val, n, nz, lnz, ic_row, ic_col, ic_val);
	return lnz;
}

struct IcSys
{
	RealSys base;
	int lnz;
	std::vector<int> l_row, l_col, u_row, u_col;
	std::vector<double> l_val, u_val, tmp;
};

static void ic_ax(void* inst, const lcg_float* x, lcg_float* y, const int n) { real_ax(&((IcSys*)inst)->base, x, y, n); }
static void ic_mx(void* inst, const lcg_float* r, lcg_float* z, const int n)
{	// z = L^-T L^-1 r
	IcSys* s = (IcSys*)inst;
	lcg_solve_lower_triangle_coo(s->l_row.data(), s->l_col.data(), s->l_val.data(), r, s->tmp.data(), n, s->lnz);
	lcg_solve_upper_triangle_coo(s->u_row.data(), s->u_col.data(), s->u_val.data(), s->tmp.data(), z, n, s->lnz);
}
static int ic_pf(void* inst, const lcg_float* m, const lcg_float conv, const lcg_para* p, const int n, const int k)
{
	IcSys* s = (IcSys*)inst;
	s->base.h.push(k, conv);
	return 0;
}

// PCG with M = L L^T from the reference's IC(0).  z_probe (nullable, n values): receives M^-1 B (one application of the
// preconditioner to the right-hand side), for checking the GPU triangular solves against the reference's.
extern "C" int lcgref_pcg_ic0(int n, const int* rp, const int* ci, const double* val, double* m, const double* B, const void* para,
	double* hist, int hist_cap, int* out, double* dout, double* z_probe)
{
	IcSys s;
	s.base.n = n; s.base.rp = rp; s.base.ci = ci; s.base.v = val; s.base.diag = nullptr;
	s.base.h.buf = hist; s.base.h.cap = hist_cap;
	const int nz = rp[n];
	std::vector<int> row((size_t)nz);
	for (int i = 0; i < n; i++) for (int k = rp[i]; k < rp[i + 1]; k++) row[(size_t)k] = i;
	s.lnz = lcgref_ic0_half(row.data(), ci, val, n, nz, nullptr, nullptr, nullptr);
	s.l_row.resize((size_t)s.lnz); s.l_col.resize((size_t)s.lnz); s.l_val.resize((size_t)s.lnz); s.tmp.resize((size_t)n);
	lcgref_ic0_half(row.data(), ci, val, n, nz, s.l_row.data(), s.l_col.data(), s.l_val.data());
	// U = L^T, row-sorted
	std::vector<int> cnt((size_t)n + 1, 0);
	for (int k = 0; k < s.lnz; k++) cnt[(size_t)s.l_col[(size_t)k] + 1]++;
	for (int i = 0; i < n; i++) cnt[(size_t)i + 1] += cnt[(size_t)i];
	s.u_row.resize((size_t)s.lnz); s.u_col.resize((size_t)s.lnz); s.u_val.resize((size_t)s.lnz);
	for (int k = 0; k < s.lnz; k++)
	{
		const int d = cnt[(size_t)s.l_col[(size_t)k]]++;
		s.u_row[(size_t)d] = s.l_col[(size_t)k]; s.u_col[(size_t)d] = s.l_row[(size_t)k]; s.u_val[(size_t)d] = s.l_val[(size_t)k];
	}
	if (z_probe) ic_mx(&s, B, z_probe, n);
	double t0 = omp_get_wtime();
	int ret = lcg_solver_preconditioned(ic_ax, ic_mx, ic_pf, m, B, n, (const lcg_para*)para, &s, LCG_PCG);
	double t1 = omp_get_wtime();
	if (out) { out[0] = s.base.h.last_k; out[1] = s.base.h.calls; }
	if (dout) { dout[0] = s.base.h.last_res; dout[1] = t1 - t0; }
	return ret;
}
