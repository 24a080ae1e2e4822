/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C99) of the iteration loops of liblcg's CPU solvers, used ONLY as the parity
 * checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).  Nothing under
 * liblcg_b200/ may include, link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every solver here bit-for-bit (return code,
 * iteration count, residual history, solution) against the unmodified reference built into
 * oracle/_ref/liblcg_ref.so (when that library is present) and against the golden fixtures in
 * tests/golden/ that were generated from it (oracle/make_golden.py).  The one exception is complex PCG
 * (`oc_pcg`): the reference has no buildable CPU version (Eigen-only, clcg_eigen.cpp:577-683), so it is
 * pinned indirectly — with M = I it must reproduce clbicg_symmetric (clcg.cpp:228-364) bit-for-bit.
 *
 * Each function cites the reference file:line (under /root/reference/src/lib) whose arithmetic it restates.
 * Expression shapes (operand order, where the division happens) follow the reference so that the results
 * are bitwise equal when both are built with the same flags (-O3, no -ffast-math, baseline x86-64).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex zc;

/* lcg_para / clcg_para: util.h:95-148, util.h:247-273 (same field order => same ABI layout) */
typedef struct { int max_iterations; double epsilon; int abs_diff; double restart_epsilon; double step; double sigma; double beta; int maxi_m; } o_para;
typedef struct { int max_iterations; double epsilon; int abs_diff; } o_cpara;

/* return codes: util.h:69-90 and util.h:226-242 */
enum { O_CONVERGENCE = 0, O_STOP = 1, O_ALREADY = 2,
       O_UNKNOWN = -1024, O_BAD_SIZE = -1023, O_BAD_MAXIT = -1022, O_BAD_EPS = -1021, O_BAD_RESTART = -1020,
       O_MAXIT = -1019, O_NULL_PRECOND = -1018, O_NAN = -1017, O_BAD_PTR = -1016, O_BAD_LAMBDA = -1015,
       O_BAD_SIGMA = -1014, O_BAD_BETA = -1013, O_BAD_MAXIM = -1012,
       OC_BAD_SIZE = -1023, OC_BAD_MAXIT = -1022, OC_BAD_EPS = -1021, OC_MAXIT_ALIAS = -1019 /* the complex solvers
       return LCG_REACHED_MAX_ITERATIONS (clcg.cpp:164), numerically CLCG_NAN_VALUE */, OC_NAN = -1019, OC_BAD_PTR = -1018 };

/* solver ids: util.h:32-64, util.h:187-221 */
enum { S_CG = 0, S_PCG, S_CGS, S_BICGSTAB, S_BICGSTAB2, S_PG, S_SPG };
enum { C_BICG = 0, C_BICG_SYM, C_CGS, C_BICGSTAB, C_TFQMR, C_PCG, C_PBICG };

typedef struct {
	int last_k, calls, cap, stop_at;
	double last_res;
	double* buf;
} o_hist;

static int hist_push(o_hist* h, int k, double r)
{
	if (!h) return 0;
	h->last_k = k; h->last_res = r;
	if (h->buf && h->calls < h->cap) h->buf[h->calls] = r;
	h->calls++;
	return (h->stop_at >= 0 && k == h->stop_at) ? 1 : 0;
}

static long g_seed_time = 0;
void lcgoracle_set_time(long t) { g_seed_time = t; }

/* Summation order of the inner products.  0 (default) = left-to-right, the reference's order (algebra.cpp:154-163,
 * lcg_complex.cpp:143-167) — the only order the golden vectors and the bit-exact pinning use.  1 = pairwise (tree), the
 * order class of the GPU's reductions: used by the GPU parity tests to tell "same algorithm, more accurate sums" (which
 * shortens the reference's erratic complex BiCG runs by ~5 %) from a real discrepancy. */
static int g_sum_tree = 0;
void lcgoracle_set_summation(int tree) { g_sum_tree = tree ? 1 : 0; }

/* ---------------------------------------------------------------- real helpers */
typedef struct { int n; const int* rp; const int* ci; const double* v; const double* diag; } o_csr;

static void o_ax(const o_csr* A, const double* x, double* y)
{
	int i;
#pragma omp parallel for schedule(static)
	for (i = 0; i < A->n; i++)
	{
		double acc = 0.0;
		for (int k = A->rp[i]; k < A->rp[i + 1]; k++) acc += A->v[k] * x[A->ci[k]];
		y[i] = acc;
	}
}

static void o_mx(const o_csr* A, const double* x, double* y)
{
	int i;
#pragma omp parallel for schedule(static)
	for (i = 0; i < A->n; i++) y[i] = x[i] / A->diag[i];
}

/* algebra.cpp:154-163 — serial left-to-right accumulation */
static double o_dot_tree(const double* a, const double* b, int lo, int hi)
{
	if (hi - lo <= 8) { double s = 0.0; for (int i = lo; i < hi; i++) s += a[i] * b[i]; return s; }
	int mid = lo + (hi - lo) / 2;
	return o_dot_tree(a, b, lo, mid) + o_dot_tree(a, b, mid, hi);
}

static double o_dot(const double* a, const double* b, int n)
{
	if (g_sum_tree) return o_dot_tree(a, b, 0, n);
	double s = 0.0;
	for (int i = 0; i < n; i++) s += a[i] * b[i];
	return s;
}

static int o_has_nan(const double* a, int n)
{
	for (int i = 0; i < n; i++) if (a[i] != a[i]) return 1;
	return 0;
}

/* algebra.cpp:50-58 with closed bounds (the only way the solvers call it, lcg.cpp:1089) */
static double o_box(double lo, double hi, double a)
{
	if (a >= hi) return hi;
	if (a <= lo) return lo;
	return a;
}

/* The loop head shared by every real solver (lcg.cpp:206-230, 361-385, 520-544, 692-716, 877-901, 1128-1152, 1318-1342).
 * Returns 1 when the loop must end (ret filled in), 0 after `t` has been incremented. */
static int o_head(const o_para* p, o_hist* h, int use_pf, double sq_res, double m_mod, int n, int* t, int* ret)
{
	double residual = p->abs_diff ? sqrt(sq_res) / n : sq_res / m_mod;
	if (use_pf && hist_push(h, *t, residual)) { *ret = O_STOP; return 1; }
	if (residual <= p->epsilon) { *ret = O_CONVERGENCE; return 1; }
	if (p->max_iterations > 0 && *t + 1 > p->max_iterations) { *ret = O_MAXIT; return 1; }
	(*t)++;
	return 0;
}

/* The "already optimised" test that precedes every loop (lcg.cpp:185-203 and clones). */
static int o_already(const o_para* p, o_hist* h, int use_pf, double sq_res, double m_mod, int n)
{
	if (p->abs_diff && sqrt(sq_res) / n <= p->epsilon)
	{
		if (use_pf) hist_push(h, 0, sqrt(sq_res) / n);
		return 1;
	}
	else if (sq_res / m_mod <= p->epsilon)
	{
		if (use_pf) hist_push(h, 0, sq_res / m_mod);
		return 1;
	}
	return 0;
}

static int o_check_common(int n, const o_para* p, const double* m, const double* B)
{
	/* lcg.cpp:150-155 */
	if (n <= 0) return O_BAD_SIZE;
	if (p->max_iterations < 0) return O_BAD_MAXIT;
	if (p->epsilon <= 0.0 || p->epsilon >= 1.0) return O_BAD_EPS;
	if (!m || !B) return O_BAD_PTR;
	return 0;
}

#define VEC(name) double* name = (double*)malloc(sizeof(double) * (size_t)n)

/* lcg.cpp:143-274 — CG in the reference's sign convention g = Ax - B, d = -g */
static int o_cg(const o_csr* A, double* m, const double* B, int n, const o_para* p, o_hist* h, int pf)
{
	int ret = o_check_common(n, p, m, B); if (ret) return ret;
	VEC(g); VEC(d); VEC(Ad);
	int i, t = 0;
	o_ax(A, m, Ad);
	for (i = 0; i < n; i++) { g[i] = Ad[i] - B[i]; d[i] = -1.0 * g[i]; }
	double m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
	double g_mod = o_dot(g, g, n);
	if (o_already(p, h, pf, g_mod, m_mod, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (o_head(p, h, pf, g_mod, m_mod, n, &t, &ret)) break;
		o_ax(A, d, Ad);
		double dAd = o_dot(d, Ad, n);
		double ak = g_mod / dAd;
		for (i = 0; i < n; i++) { m[i] += ak * d[i]; g[i] += ak * Ad[i]; }
		m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
		if (o_has_nan(m, n)) { ret = O_NAN; break; }
		double g1 = o_dot(g, g, n);
		double bk = g1 / g_mod; g_mod = g1;
		for (i = 0; i < n; i++) d[i] = bk * d[i] - g[i];
	}
done:
	free(g); free(d); free(Ad);
	return ret;
}

/* lcg.cpp:293-434 — PCG (Kaasschieter alg. 1); convergence still on r.r / max(m.m,1) */
static int o_pcg(const o_csr* A, double* m, const double* B, int n, const o_para* p, o_hist* h, int pf)
{
	int ret = o_check_common(n, p, m, B); if (ret) return ret;
	VEC(r); VEC(z); VEC(d); VEC(Ad);
	int i, t = 0;
	o_ax(A, m, Ad);
	for (i = 0; i < n; i++) r[i] = B[i] - Ad[i];
	o_mx(A, r, z);
	for (i = 0; i < n; i++) d[i] = z[i];
	double m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
	double r_mod = o_dot(r, r, n);
	double zr = o_dot(z, r, n);
	if (o_already(p, h, pf, r_mod, m_mod, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (o_head(p, h, pf, r_mod, m_mod, n, &t, &ret)) break;
		o_ax(A, d, Ad);
		double dAd = o_dot(d, Ad, n);
		double ak = zr / dAd;
		for (i = 0; i < n; i++) { m[i] += ak * d[i]; r[i] -= ak * Ad[i]; }
		o_mx(A, r, z);
		m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
		r_mod = o_dot(r, r, n);
		if (o_has_nan(m, n)) { ret = O_NAN; break; }
		double zr1 = o_dot(z, r, n);
		double bk = zr1 / zr; zr = zr1;
		for (i = 0; i < n; i++) d[i] = z[i] + bk * d[i];
	}
done:
	free(r); free(z); free(d); free(Ad);
	return ret;
}

/* lcg.cpp:437-612 — CGS with shadow residual r0~ = r0 */
static int o_cgs(const o_csr* A, double* m, const double* B, int n, const o_para* p, o_hist* h, int pf)
{
	int ret = o_check_common(n, p, m, B); if (ret) return ret;
	VEC(r); VEC(r0); VEC(pk); VEC(Ax); VEC(u); VEC(q); VEC(w);
	int i, t = 0;
	o_ax(A, m, Ax);
	for (i = 0; i < n; i++) pk[i] = u[i] = r0[i] = r[i] = B[i] - Ax[i];
	double rr0 = o_dot(r, r0, n);
	double m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
	double r_mod = o_dot(r, r, n);
	if (o_already(p, h, pf, r_mod, m_mod, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (o_head(p, h, pf, r_mod, m_mod, n, &t, &ret)) break;
		o_ax(A, pk, Ax);
		double Apr = o_dot(Ax, r0, n);
		double ak = rr0 / Apr;
		for (i = 0; i < n; i++) { q[i] = u[i] - ak * Ax[i]; w[i] = u[i] + q[i]; }
		o_ax(A, w, Ax);
		for (i = 0; i < n; i++) { m[i] += ak * w[i]; r[i] -= ak * Ax[i]; }
		m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
		r_mod = o_dot(r, r, n);
		if (o_has_nan(m, n)) { ret = O_NAN; break; }
		double rr1 = o_dot(r, r0, n);
		double bk = rr1 / rr0; rr0 = rr1;
		for (i = 0; i < n; i++) { u[i] = r[i] + bk * q[i]; pk[i] = u[i] + bk * (q[i] + bk * pk[i]); }
	}
done:
	free(r); free(r0); free(pk); free(Ax); free(u); free(q); free(w);
	return ret;
}

/* lcg.cpp:629-794 (restart = 0) and lcg.cpp:812-1034 (restart = 1, "BICGSTAB2") */
static int o_bicgstab(const o_csr* A, double* m, const double* B, int n, const o_para* p, o_hist* h, int pf, int restart)
{
	int ret;
	if (!restart) { ret = o_check_common(n, p, m, B); if (ret) return ret; }
	else
	{	/* lcg.cpp:819-825 — note the oddly placed epsilon >= 1 test */
		if (n <= 0) return O_BAD_SIZE;
		if (p->max_iterations < 0) return O_BAD_MAXIT;
		if (p->epsilon <= 0.0) return O_BAD_EPS;
		if (p->restart_epsilon <= 0.0 || p->epsilon >= 1.0) return O_BAD_RESTART;
		if (!m || !B) return O_BAD_PTR;
	}
	VEC(r); VEC(r0); VEC(pk); VEC(Ax); VEC(s); VEC(Ap);
	int i, t = 0;
	o_ax(A, m, Ax);
	for (i = 0; i < n; i++) pk[i] = r0[i] = r[i] = B[i] - Ax[i];
	double rr0 = o_dot(r, r0, n);
	double m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
	double r_mod = o_dot(r, r, n);
	if (o_already(p, h, pf, r_mod, m_mod, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (o_head(p, h, pf, r_mod, m_mod, n, &t, &ret)) break;
		o_ax(A, pk, Ap);
		double Apr = o_dot(Ap, r0, n);
		double ak = rr0 / Apr;
		for (i = 0; i < n; i++) s[i] = r[i] - ak * Ap[i];
		if (restart && p->abs_diff)
		{	/* lcg.cpp:918-950 — half-step test on s; t is incremented a second time */
			double res = sqrt(o_dot(s, s, n)) / n;
			if (pf && hist_push(h, t, res)) { ret = O_STOP; break; }
			if (res <= p->epsilon)
			{
				int nan = 0;
				for (i = 0; i < n; i++) { m[i] += ak * pk[i]; if (m[i] != m[i]) { nan = 1; break; } }
				ret = nan ? O_NAN : O_CONVERGENCE; break;
			}
			if (p->max_iterations > 0 && t + 1 > p->max_iterations) { ret = O_MAXIT; break; }
			t++;
		}
		o_ax(A, s, Ax);
		double Ass = 0.0, AsAs = 0.0;
		for (i = 0; i < n; i++) { Ass += Ax[i] * s[i]; AsAs += Ax[i] * Ax[i]; }
		double wk = Ass / AsAs;
		if (!restart) for (i = 0; i < n; i++) m[i] += (ak * pk[i] + wk * s[i]);
		else          for (i = 0; i < n; i++) m[i] += ak * pk[i] + wk * s[i];
		m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
		if (o_has_nan(m, n)) { ret = O_NAN; break; }
		for (i = 0; i < n; i++) r[i] = s[i] - wk * Ax[i];
		r_mod = o_dot(r, r, n);
		double rr1 = o_dot(r, r0, n);
		if (restart && fabs(rr1) < p->restart_epsilon)
		{	/* lcg.cpp:993-1009 */
			for (i = 0; i < n; i++) { r0[i] = r[i]; pk[i] = r[i]; }
			rr1 = o_dot(r, r0, n);
			rr0 = rr1;	/* betak is computed but unused on this branch */
		}
		else
		{
			double bk = (ak / wk) * rr1 / rr0; rr0 = rr1;
			for (i = 0; i < n; i++) pk[i] = r[i] + bk * (pk[i] - wk * Ap[i]);
		}
	}
done:
	free(r); free(r0); free(pk); free(Ax); free(s); free(Ap);
	return ret;
}

/* lcg.cpp:1054-1204 — projected gradient with Barzilai-Borwein step */
static int o_pg(const o_csr* A, double* m, const double* B, const double* lo, const double* hi, int n,
	const o_para* p, o_hist* h, int pf)
{
	/* lcg.cpp:1062-1070 */
	if (n <= 0) return O_BAD_SIZE;
	if (p->max_iterations < 0) return O_BAD_MAXIT;
	if (p->epsilon <= 0.0) return O_BAD_EPS;
	if (p->step <= 0.0 || p->epsilon >= 1.0) return O_BAD_LAMBDA;
	if (!m || !B || !lo || !hi) return O_BAD_PTR;
	int ret, i, t = 0;
	VEC(g); VEC(Ad); VEC(mn); VEC(gn); VEC(s); VEC(y);
	double alpha = p->step;
	for (i = 0; i < n; i++) m[i] = o_box(lo[i], hi[i], m[i]);
	o_ax(A, m, Ad);
	for (i = 0; i < n; i++) g[i] = Ad[i] - B[i];
	double m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
	double g_mod = o_dot(g, g, n);
	if (o_already(p, h, pf, g_mod, m_mod, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (o_head(p, h, pf, g_mod, m_mod, n, &t, &ret)) break;
		for (i = 0; i < n; i++) mn[i] = o_box(lo[i], hi[i], m[i] - alpha * g[i]);
		o_ax(A, mn, Ad);
		for (i = 0; i < n; i++) { gn[i] = Ad[i] - B[i]; s[i] = mn[i] - m[i]; y[i] = gn[i] - g[i]; }
		double ss = 0.0, sy = 0.0;
		for (i = 0; i < n; i++) { ss += s[i] * s[i]; sy += s[i] * y[i]; }
		alpha = ss / sy;
		for (i = 0; i < n; i++) { m[i] = mn[i]; g[i] = gn[i]; }
		m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
		g_mod = o_dot(g, g, n);
	}
done:
	free(g); free(Ad); free(mn); free(gn); free(s); free(y);
	return ret;
}

/* lcg.cpp:1224-1447 — spectral projected gradient with non-monotone line search */
static int o_spg(const o_csr* A, double* m, const double* B, const double* lo, const double* hi, int n,
	const o_para* p, o_hist* h, int pf)
{
	/* lcg.cpp:1232-1243 */
	if (n <= 0) return O_BAD_SIZE;
	if (p->max_iterations < 0) return O_BAD_MAXIT;
	if (p->epsilon <= 0.0 || p->epsilon >= 1.0) return O_BAD_EPS;
	if (p->step <= 0.0) return O_BAD_LAMBDA;
	if (p->sigma <= 0.0 || p->sigma >= 1.0) return O_BAD_SIGMA;
	if (p->beta <= 0.0 || p->beta >= 1.0) return O_BAD_BETA;
	if (p->maxi_m <= 0) return O_BAD_MAXIM;
	if (!m || !B || !lo || !hi) return O_BAD_PTR;
	int ret, i, t = 0;
	VEC(g); VEC(Ad); VEC(mn); VEC(gn); VEC(s); VEC(y); VEC(d);
	double* qm = (double*)malloc(sizeof(double) * (size_t)p->maxi_m);
	double lambda = p->step, qk = 0;
	for (i = 0; i < n; i++) m[i] = o_box(lo[i], hi[i], m[i]);
	o_ax(A, m, Ad);
	for (i = 0; i < n; i++) g[i] = Ad[i] - B[i];
	double m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
	double g_mod = o_dot(g, g, n);
	if (o_already(p, h, pf, g_mod, m_mod, n)) { ret = O_ALREADY; goto done; }
	for (i = 0; i < n; i++) qk += (0.5 * m[i] * Ad[i] - B[i] * m[i]);
	qm[0] = qk;
	for (i = 1; i < p->maxi_m; i++) qm[i] = -1e+30;
	while (1)
	{
		if (o_head(p, h, pf, g_mod, m_mod, n, &t, &ret)) break;
		for (i = 0; i < n; i++) d[i] = o_box(lo[i], hi[i], m[i] - lambda * g[i]) - m[i];
		double alpha = 1.0;
		while (1)
		{	/* first trial lcg.cpp:1351-1369, shrinking trials lcg.cpp:1377-1399 */
			for (i = 0; i < n; i++) mn[i] = m[i] + alpha * d[i];
			o_ax(A, mn, Ad);
			qk = 0.0;
			for (i = 0; i < n; i++) qk += (0.5 * mn[i] * Ad[i] - B[i] * mn[i]);
			double amod = 0.0;
			for (i = 0; i < n; i++) amod += p->sigma * alpha * g[i] * d[i];
			double qmax = qm[0];
			for (i = 1; i < p->maxi_m; i++) qmax = (qmax >= qm[i]) ? qmax : qm[i];
			if (!(qk > qmax + amod)) break;
			alpha = alpha * p->beta;
		}
		qm[(t + 1) % p->maxi_m] = qk;
		for (i = 0; i < n; i++) { gn[i] = Ad[i] - B[i]; s[i] = mn[i] - m[i]; y[i] = gn[i] - g[i]; }
		double ss = 0.0, sy = 0.0;
		for (i = 0; i < n; i++) { ss += s[i] * s[i]; sy += s[i] * y[i]; }
		lambda = ss / sy;
		for (i = 0; i < n; i++) { m[i] = mn[i]; g[i] = gn[i]; }
		m_mod = o_dot(m, m, n); if (m_mod < 1.0) m_mod = 1.0;
		g_mod = o_dot(g, g, n);
	}
done:
	free(g); free(Ad); free(mn); free(gn); free(s); free(y); free(d); free(qm);
	return ret;
}

static void fill_out(o_hist* h, int* out, double* dout, double secs)
{
	if (out) { out[0] = h->last_k; out[1] = h->calls; }
	if (dout) { dout[0] = h->last_res; dout[1] = secs; }
}

static double now_s(void)
{
#ifdef _OPENMP
	return omp_get_wtime();
#else
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

/* Same signature as lcgref_solve in oracle/ref_shim.cpp.  Dispatch mirrors lcg.cpp:59-82, 87-91, 121-140. */
int lcgoracle_solve(int solver_id, int n, const int* rp, const int* ci, const double* val,
	double* m, const double* B, const double* low, const double* hig, const double* diag,
	const void* para, int use_progress, int stop_at, double* hist, int hist_cap, int* out, double* dout)
{
	o_para defp = { 0, 1e-6, 0, 1e-6, 1.0, 0.95, 0.9, 10 };	/* util.h:153 */
	o_para p = para ? *(const o_para*)para : defp;
	o_csr A = { n, rp, ci, val, diag };
	o_hist h = { -1, 0, hist_cap, stop_at, 0.0, hist };
	int ret;
	double t0 = now_s();
	switch (solver_id)
	{
		case S_CG: ret = o_cg(&A, m, B, n, &p, &h, use_progress); break;
		case S_PCG: ret = o_pcg(&A, m, B, n, &p, &h, use_progress); break;
		case S_BICGSTAB: ret = o_bicgstab(&A, m, B, n, &p, &h, use_progress, 0); break;
		case S_BICGSTAB2: ret = o_bicgstab(&A, m, B, n, &p, &h, use_progress, 1); break;
		case S_PG: ret = o_pg(&A, m, B, low, hig, n, &p, &h, use_progress); break;
		case S_SPG: ret = o_spg(&A, m, B, low, hig, n, &p, &h, use_progress); break;
		case S_CGS: default: ret = o_cgs(&A, m, B, n, &p, &h, use_progress); break;
	}
	fill_out(&h, out, dout, now_s() - t0);
	return ret;
}

void lcgoracle_spmv(int n, const int* rp, const int* ci, const double* val, const double* x, double* y)
{
	o_csr A = { n, rp, ci, val, 0 };
	o_ax(&A, x, y);
}

/* ================================================================ complex */
typedef struct { int n; const int* rp; const int* ci; const zc* v; int* trp; int* tci; zc* tv; const zc* diag; } oc_csr;

static void oc_transpose(oc_csr* A)
{
	int n = A->n, nnz = A->rp[n], i, k;
	A->trp = (int*)calloc((size_t)n + 1, sizeof(int));
	A->tci = (int*)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
	A->tv = (zc*)malloc(sizeof(zc) * (size_t)(nnz > 0 ? nnz : 1));
	int* fill = (int*)malloc(sizeof(int) * (size_t)n);
	for (k = 0; k < nnz; k++) A->trp[A->ci[k] + 1]++;
	for (i = 0; i < n; i++) A->trp[i + 1] += A->trp[i];
	for (i = 0; i < n; i++) fill[i] = A->trp[i];
	for (i = 0; i < n; i++)
		for (k = A->rp[i]; k < A->rp[i + 1]; k++) { int d = fill[A->ci[k]]++; A->tci[d] = i; A->tv[d] = A->v[k]; }
	free(fill);
}

/* op: 0 = A x, 1 = A^T x, 2 = A^H x, 3 = conj(A) x  (clcg.h:40 layout/conjugate pairs) */
static void oc_ax(const oc_csr* A, const zc* x, zc* y, int op)
{
	const int tr = (op == 1 || op == 2), cj = (op == 2 || op == 3);
	const int* rp = tr ? A->trp : A->rp; const int* ci = tr ? A->tci : A->ci; const zc* v = tr ? A->tv : A->v;
	int i;
#pragma omp parallel for schedule(static)
	for (i = 0; i < A->n; i++)
	{
		zc acc = 0.0;
		for (int k = rp[i]; k < rp[i + 1]; k++) acc += (cj ? conj(v[k]) : v[k]) * x[ci[k]];
		y[i] = acc;
	}
}

/* lcg_complex.cpp:156-167 — <a,b> = sum conj(a_i) b_i, serial */
static zc oc_sum_tree(const zc* a, const zc* b, int lo, int hi, int conj_first)
{
	if (hi - lo <= 8)
	{
		double re = 0.0, im = 0.0;
		for (int i = lo; i < hi; i++)
		{
			if (conj_first) { re += (creal(a[i]) * creal(b[i]) + cimag(a[i]) * cimag(b[i])); im += (creal(a[i]) * cimag(b[i]) - cimag(a[i]) * creal(b[i])); }
			else { re += (creal(a[i]) * creal(b[i]) - cimag(a[i]) * cimag(b[i])); im += (creal(a[i]) * cimag(b[i]) + cimag(a[i]) * creal(b[i])); }
		}
		return CMPLX(re, im);
	}
	int mid = lo + (hi - lo) / 2;
	return oc_sum_tree(a, b, lo, mid, conj_first) + oc_sum_tree(a, b, mid, hi, conj_first);
}

static zc oc_inner(const zc* a, const zc* b, int n)
{
	if (g_sum_tree) return oc_sum_tree(a, b, 0, n, 1);
	double re = 0.0, im = 0.0;
	for (int i = 0; i < n; i++)
	{
		re += (creal(a[i]) * creal(b[i]) + cimag(a[i]) * cimag(b[i]));
		im += (creal(a[i]) * cimag(b[i]) - cimag(a[i]) * creal(b[i]));
	}
	return CMPLX(re, im);
}

/* lcg_complex.cpp:143-154 — unconjugated sum a_i b_i, serial */
static zc oc_dot(const zc* a, const zc* b, int n)
{
	if (g_sum_tree) return oc_sum_tree(a, b, 0, n, 0);
	double re = 0.0, im = 0.0;
	for (int i = 0; i < n; i++)
	{
		re += (creal(a[i]) * creal(b[i]) - cimag(a[i]) * cimag(b[i]));
		im += (creal(a[i]) * cimag(b[i]) + cimag(a[i]) * creal(b[i]));
	}
	return CMPLX(re, im);
}

static double oc_square(zc a) { return creal(a) * creal(a) + cimag(a) * cimag(a); }	/* lcg_complex.cpp:102-105 */
static double oc_module(zc a) { return sqrt(oc_square(a)); }				/* lcg_complex.cpp:107-110 */

static int oc_has_nan(const zc* a, int n)
{
	/* `m[i] != m[i]` on std::complex is true when either part is NaN */
	for (int i = 0; i < n; i++) if (creal(a[i]) != creal(a[i]) || cimag(a[i]) != cimag(a[i])) return 1;
	return 0;
}

/* lcg_complex.cpp:118-127 with l = 1+0i, h = 2+0i; the imaginary part still consumes one rand() */
static void oc_vecrnd(zc* a, int n)
{
	srand((unsigned)(g_seed_time ? g_seed_time : time(0)));
	for (int i = 0; i < n; i++)
	{
		double re = (2.0 - 1.0) * rand() * 1.0 / RAND_MAX + 1.0;
		double im = (0.0 - 0.0) * rand() * 1.0 / RAND_MAX + 0.0;
		a[i] = CMPLX(re, im);
	}
}

void lcgoracle_vecrnd(double* a, int n) { oc_vecrnd((zc*)a, n); }

static int oc_check(int n, const o_cpara* p, const zc* m, const zc* B)
{
	if (n <= 0) return OC_BAD_SIZE;
	if (p->max_iterations < 0) return OC_BAD_MAXIT;
	if (p->epsilon <= 0.0 || p->epsilon >= 1.0) return OC_BAD_EPS;
	if (!m || !B) return OC_BAD_PTR;
	return 0;
}

/* complex loop head: clcg.cpp:144-168 and clones.  rk_square = |<r,r>|^2 = ||r||^4, m_square = max(||m||^4, 1) */
static int oc_head(const o_cpara* p, o_hist* h, int pf, double rk_square, double m_square, int n, int* t, int* ret)
{
	double residual = p->abs_diff ? sqrt(rk_square) / n : rk_square / m_square;
	if (pf && hist_push(h, *t, residual)) { *ret = O_STOP; return 1; }
	if (residual <= p->epsilon) { *ret = O_CONVERGENCE; return 1; }
	if (p->max_iterations > 0 && *t + 1 > p->max_iterations) { *ret = OC_MAXIT_ALIAS; return 1; }
	(*t)++;
	return 0;
}

static int oc_already(const o_cpara* p, o_hist* h, int pf, double rk_square, double m_square, int n)
{
	if (p->abs_diff && sqrt(rk_square) / n <= p->epsilon) { if (pf) hist_push(h, 0, sqrt(rk_square) / n); return 1; }
	else if (rk_square / m_square <= p->epsilon) { if (pf) hist_push(h, 0, rk_square / m_square); return 1; }
	return 0;
}

static double oc_msq(const zc* m, int n) { double s = oc_square(oc_inner(m, m, n)); return s < 1.0 ? 1.0 : s; }

#define ZVEC(name) zc* name = (zc*)malloc(sizeof(zc) * (size_t)n)

/* clcg.cpp:77-226 — BiCG */
static int oc_bicg(const oc_csr* A, zc* m, const zc* B, int n, const o_cpara* p, o_hist* h, int pf)
{
	int ret = oc_check(n, p, m, B); if (ret) return ret;
	ZVEC(r1); ZVEC(r2); ZVEC(d1); ZVEC(d2); ZVEC(Ax);
	int i, t = 0;
	oc_ax(A, m, Ax, 0);
	for (i = 0; i < n; i++) { d1[i] = r1[i] = B[i] - Ax[i]; d2[i] = r2[i] = conj(r1[i]); }
	zc r1r2 = oc_inner(r2, r1, n);
	double msq = oc_msq(m, n);
	double rsq = oc_square(oc_inner(r1, r1, n));
	if (oc_already(p, h, pf, rsq, msq, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (oc_head(p, h, pf, rsq, msq, n, &t, &ret)) break;
		oc_ax(A, d1, Ax, 0);
		zc Ad1d2 = oc_inner(d2, Ax, n);
		zc ak = r1r2 / Ad1d2;
		for (i = 0; i < n; i++) { m[i] = m[i] + ak * d1[i]; r1[i] = r1[i] - ak * Ax[i]; }
		msq = oc_msq(m, n);
		rsq = oc_square(oc_inner(r1, r1, n));
		oc_ax(A, d2, Ax, 2);
		for (i = 0; i < n; i++) r2[i] = r2[i] - conj(ak) * Ax[i];
		if (oc_has_nan(m, n)) { ret = OC_NAN; break; }
		zc nxt = oc_inner(r2, r1, n);
		zc bk = nxt / r1r2; r1r2 = nxt;
		for (i = 0; i < n; i++) { d1[i] = r1[i] + bk * d1[i]; d2[i] = r2[i] + conj(bk) * d2[i]; }
	}
done:
	free(r1); free(r2); free(d1); free(d2); free(Ax);
	return ret;
}

/* clcg.cpp:228-364 (precond = 0) — BiCG for complex symmetric A == CG with the unconjugated dot;
 * precond = 1 — complex PCG following clcg_eigen.cpp:577-683 / clcg_cuda.cu:403-559 (unconjugated r.z). */
static int oc_cgsym(const oc_csr* A, zc* m, const zc* B, int n, const o_cpara* p, o_hist* h, int pf, int precond)
{
	int ret = oc_check(n, p, m, B); if (ret) return ret;
	ZVEC(r); ZVEC(d); ZVEC(Ax); ZVEC(s);
	int i, t = 0;
	oc_ax(A, m, Ax, 0);
	if (!precond) for (i = 0; i < n; i++) d[i] = r[i] = B[i] - Ax[i];
	else { for (i = 0; i < n; i++) r[i] = B[i] - Ax[i]; for (i = 0; i < n; i++) d[i] = r[i] / A->diag[i]; }
	zc rho = oc_dot(r, d, n);	/* == r.r when not preconditioned */
	double msq = oc_msq(m, n);
	double rsq = oc_square(oc_inner(r, r, n));
	if (oc_already(p, h, pf, rsq, msq, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (oc_head(p, h, pf, rsq, msq, n, &t, &ret)) break;
		oc_ax(A, d, Ax, 0);
		zc dAx = oc_dot(d, Ax, n);
		zc ak = rho / dAx;
		for (i = 0; i < n; i++) { m[i] = m[i] + ak * d[i]; r[i] = r[i] - ak * Ax[i]; }
		msq = oc_msq(m, n);
		rsq = oc_square(oc_inner(r, r, n));
		if (!precond && oc_has_nan(m, n)) { ret = OC_NAN; break; }
		zc rho2;
		if (!precond) rho2 = oc_dot(r, r, n);
		else { for (i = 0; i < n; i++) s[i] = r[i] / A->diag[i]; rho2 = oc_dot(r, s, n); }
		zc bk = rho2 / rho; rho = rho2;
		if (!precond) for (i = 0; i < n; i++) d[i] = r[i] + bk * d[i];
		else          for (i = 0; i < n; i++) d[i] = s[i] + bk * d[i];
	}
done:
	free(r); free(d); free(Ax); free(s);
	return ret;
}

/* clcg.cpp:366-522 — CGS with a random shadow residual */
static int oc_cgs(const oc_csr* A, zc* m, const zc* B, int n, const o_cpara* p, o_hist* h, int pf)
{
	int ret = oc_check(n, p, m, B); if (ret) return ret;
	ZVEC(r); ZVEC(rb); ZVEC(pk); ZVEC(Ax); ZVEC(u); ZVEC(q); ZVEC(w);
	int i, t = 0;
	oc_ax(A, m, Ax, 0);
	for (i = 0; i < n; i++) pk[i] = u[i] = r[i] = B[i] - Ax[i];
	zc rho;
	do { oc_vecrnd(rb, n); rho = oc_inner(rb, r, n); } while (oc_module(rho) < 1e-8);
	double msq = oc_msq(m, n);
	double rsq = oc_square(oc_inner(r, r, n));
	if (oc_already(p, h, pf, rsq, msq, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (oc_head(p, h, pf, rsq, msq, n, &t, &ret)) break;
		oc_ax(A, pk, Ax, 0);
		zc sigma = oc_inner(rb, Ax, n);
		zc ak = rho / sigma;
		for (i = 0; i < n; i++) { q[i] = u[i] - ak * Ax[i]; w[i] = u[i] + q[i]; }
		oc_ax(A, w, Ax, 0);
		for (i = 0; i < n; i++) { m[i] = m[i] + ak * w[i]; r[i] = r[i] - ak * Ax[i]; }
		msq = oc_msq(m, n);
		rsq = oc_square(oc_inner(r, r, n));
		if (oc_has_nan(m, n)) { ret = OC_NAN; break; }
		zc rho2 = oc_inner(rb, r, n);
		zc bk = rho2 / rho; rho = rho2;
		for (i = 0; i < n; i++) { u[i] = r[i] + bk * q[i]; pk[i] = u[i] + bk * (q[i] + bk * pk[i]); }
	}
done:
	free(r); free(rb); free(pk); free(Ax); free(u); free(q); free(w);
	return ret;
}

/* clcg.cpp:524-679 — BiCGSTAB */
static int oc_bicgstab(const oc_csr* A, zc* m, const zc* B, int n, const o_cpara* p, o_hist* h, int pf)
{
	int ret = oc_check(n, p, m, B); if (ret) return ret;
	ZVEC(r); ZVEC(rb); ZVEC(pk); ZVEC(s); ZVEC(Ap); ZVEC(As);
	int i, t = 0;
	oc_ax(A, m, Ap, 0);
	for (i = 0; i < n; i++) pk[i] = r[i] = B[i] - Ap[i];
	zc rho;
	do { oc_vecrnd(rb, n); rho = oc_inner(rb, r, n); } while (oc_module(rho) < 1e-8);
	double msq = oc_msq(m, n);
	double rsq = oc_square(oc_inner(r, r, n));
	if (oc_already(p, h, pf, rsq, msq, n)) { ret = O_ALREADY; goto done; }
	while (1)
	{
		if (oc_head(p, h, pf, rsq, msq, n, &t, &ret)) break;
		oc_ax(A, pk, Ap, 0);
		zc sigma = oc_inner(rb, Ap, n);
		zc ak = rho / sigma;
		for (i = 0; i < n; i++) s[i] = r[i] - ak * Ap[i];
		oc_ax(A, s, As, 0);
		zc Ass = oc_inner(As, s, n);
		zc AsAs = oc_inner(As, As, n);
		zc omega = Ass / AsAs;
		for (i = 0; i < n; i++) { m[i] = m[i] + ak * pk[i] + omega * s[i]; r[i] = s[i] - omega * As[i]; }
		msq = oc_msq(m, n);
		rsq = oc_square(oc_inner(r, r, n));
		if (oc_has_nan(m, n)) { ret = OC_NAN; break; }
		zc rho2 = oc_inner(rb, r, n);
		zc bk = rho2 * ak / (rho * omega); rho = rho2;
		for (i = 0; i < n; i++) pk[i] = r[i] + bk * (pk[i] - omega * Ap[i]);
	}
done:
	free(r); free(rb); free(pk); free(s); free(Ap); free(As);
	return ret;
}

/* clcg.cpp:681-882 — TFQMR.  Arithmetic as in the reference (including `omega` built from ||r||^2 where
 * the textbook has ||r||, clcg.cpp:727,812,822).  ONE DELIBERATE DIFFERENCE: at clcg.cpp:800-804 the
 * max-iteration `break` only leaves the inner `for j` loop, so the reference never terminates on
 * max_iterations; here (and in the CUDA path) it returns LCG_REACHED_MAX_ITERATIONS immediately. */
static int oc_tfqmr(const oc_csr* A, zc* m, const zc* B, int n, const o_cpara* p, o_hist* h, int pf)
{
	int ret = oc_check(n, p, m, B); if (ret) return ret;
	ZVEC(pk); ZVEC(u); ZVEC(v); ZVEC(d); ZVEC(rb); ZVEC(r); ZVEC(Ax); ZVEC(q); ZVEC(uq);
	int i, j, t = 0, stop = 0;
	oc_ax(A, m, Ax, 0);
	for (i = 0; i < n; i++) { pk[i] = u[i] = r[i] = B[i] - Ax[i]; d[i] = 0.0; }
	zc rk_mod = oc_inner(r, r, n), rk_mod2;
	double rsq = oc_square(rk_mod);
	zc rho;
	do { oc_vecrnd(rb, n); rho = oc_inner(rb, r, n); } while (oc_module(rho) < 1e-8);
	double theta = 0.0, omega = oc_module(rk_mod), tao = omega;
	zc eta = 0.0;
	double msq = oc_msq(m, n);
	if (oc_already(p, h, pf, rsq, msq, n)) { ret = O_ALREADY; goto done; }
	while (!stop)
	{
		oc_ax(A, pk, v, 0);
		zc sigma = oc_inner(rb, v, n);
		zc alpha = rho / sigma;
		for (i = 0; i < n; i++) { q[i] = u[i] - alpha * v[i]; uq[i] = u[i] + q[i]; }
		oc_ax(A, uq, Ax, 0);
		for (i = 0; i < n; i++) r[i] = r[i] - alpha * Ax[i];
		rk_mod2 = oc_inner(r, r, n);
		for (j = 1; j <= 2; j++)
		{
			if (oc_head(p, h, pf, rsq, msq, n, &t, &ret)) { stop = 1; break; }
			zc sign = theta * theta * (eta / alpha);
			if (j == 1)
			{
				omega = sqrt(oc_module(rk_mod) * oc_module(rk_mod2));
				for (i = 0; i < n; i++) d[i] = u[i] + sign * d[i];
			}
			else
			{
				omega = oc_module(rk_mod2);
				for (i = 0; i < n; i++) d[i] = q[i] + sign * d[i];
			}
			theta = omega / tao;
			tao = omega / sqrt(1.0 + theta * theta);
			eta = (1.0 / (1.0 + theta * theta)) * alpha;
			for (i = 0; i < n; i++) m[i] = m[i] + eta * d[i];
			msq = oc_msq(m, n);
			if (oc_has_nan(m, n)) { ret = OC_NAN; stop = 1; break; }
		}
		if (stop) break;
		rk_mod = rk_mod2;
		rsq = oc_square(rk_mod);
		zc rho2 = oc_inner(rb, r, n);
		zc bk = rho2 / rho; rho = rho2;
		for (i = 0; i < n; i++) { u[i] = r[i] + bk * q[i]; pk[i] = u[i] + bk * (q[i] + bk * pk[i]); }
	}
done:
	free(pk); free(u); free(v); free(d); free(rb); free(r); free(Ax); free(q); free(uq);
	return ret;
}

/* Same signature as lcgref_csolve plus `diag` (complex Jacobi diagonal, used by C_PCG only).
 * Dispatch mirrors clcg.cpp:46-74 (default CGS); C_PCG mirrors clcg_cuda.cu:70-84. */
int lcgoracle_csolve(int solver_id, int n, const int* rp, const int* ci, const double* val,
	double* m, const double* B, const double* diag, const void* para, int use_progress, int stop_at,
	double* hist, int hist_cap, int* out, double* dout)
{
	o_cpara defp = { 0, 1e-6, 0 };	/* util.h:278 */
	o_cpara p = para ? *(const o_cpara*)para : defp;
	oc_csr A = { n, rp, ci, (const zc*)val, 0, 0, 0, (const zc*)diag };
	o_hist h = { -1, 0, hist_cap, stop_at, 0.0, hist };
	int ret;
	double t0 = now_s();
	if (n > 0 && solver_id == C_BICG) oc_transpose(&A);
	switch (solver_id)
	{
		case C_BICG: ret = oc_bicg(&A, (zc*)m, (const zc*)B, n, &p, &h, use_progress); break;
		case C_BICG_SYM: ret = oc_cgsym(&A, (zc*)m, (const zc*)B, n, &p, &h, use_progress, 0); break;
		case C_BICGSTAB: ret = oc_bicgstab(&A, (zc*)m, (const zc*)B, n, &p, &h, use_progress); break;
		case C_TFQMR: ret = oc_tfqmr(&A, (zc*)m, (const zc*)B, n, &p, &h, use_progress); break;
		case C_PCG: ret = oc_cgsym(&A, (zc*)m, (const zc*)B, n, &p, &h, use_progress, 1); break;
		case C_CGS: default: ret = oc_cgs(&A, (zc*)m, (const zc*)B, n, &p, &h, use_progress); break;
	}
	free(A.trp); free(A.tci); free(A.tv);
	fill_out(&h, out, dout, now_s() - t0);
	return ret;
}

void lcgoracle_cspmv(int n, const int* rp, const int* ci, const double* val, const double* x, double* y,
	int transpose, int conjugate)
{
	oc_csr A = { n, rp, ci, (const zc*)val, 0, 0, 0, 0 };
	if (transpose) oc_transpose(&A);
	oc_ax(&A, (const zc*)x, (zc*)y, transpose ? (conjugate ? 2 : 1) : (conjugate ? 3 : 0));
	free(A.trp); free(A.tci); free(A.tv);
}

int lcgoracle_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* bench.py's CPU arms set their own thread count: a launcher (torchrun) may have exported OMP_NUM_THREADS=1 */
void lcgoracle_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void)n;
#endif
}

/* ================================================================ synthetic systems
 * Host generator of the bench workloads for the CPU arms of bench.py (cpu_baseline, --impl reference): rows
 * [row0,row1) of the g^3 7-point / 27-point Poisson or 7-point convection-diffusion matrix, columns ascending,
 * Dirichlet truncation, and b = A x* with x*[i] = uint32(i*2654435761)/2^32 summed left to right.  Same
 * definition as liblcg_b200/stencil.py (tests compare the two).  kind: 0 = 7pt, 1 = 27pt, 2 = 7pt_cd. */
static int st_row(int kind, int g, long long row, long long* cols, double* vals)
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
			cols[cnt] = ((long long)zz * g + yy) * g + xx;
			vals[cnt] = (dz == 0 && dy == 0 && dx == 0) ? 26.0 : -1.0;
			cnt++;
		}
		return cnt;
	}
	const double gx = kind == 2 ? 0.5 : 0.0, gy = kind == 2 ? 0.25 : 0.0, gz = kind == 2 ? 0.125 : 0.0;
	if (z > 0) { cols[cnt] = row - gg; vals[cnt] = -1.0 - gz; cnt++; }
	if (y > 0) { cols[cnt] = row - g; vals[cnt] = -1.0 - gy; cnt++; }
	if (x > 0) { cols[cnt] = row - 1; vals[cnt] = -1.0 - gx; cnt++; }
	cols[cnt] = row; vals[cnt] = 6.0; cnt++;
	if (x < g - 1) { cols[cnt] = row + 1; vals[cnt] = -1.0 + gx; cnt++; }
	if (y < g - 1) { cols[cnt] = row + g; vals[cnt] = -1.0 + gy; cnt++; }
	if (z < g - 1) { cols[cnt] = row + gg; vals[cnt] = -1.0 + gz; cnt++; }
	return cnt;
}

/* fills rp[0..rows] (and returns nnz); with ci/val/b non-null also the entries and the right-hand side */
long long lcgoracle_gen_stencil(int kind, int g, long long row0, long long row1, int* rp, int* ci, double* val, double* b)
{
	const long long rows = row1 - row0;
	long long cols[27]; double vals[27];
	long long acc = 0;
	for (long long r = 0; r < rows; r++) { rp[r] = (int)acc; acc += st_row(kind, g, row0 + r, cols, vals); }
	rp[rows] = (int)acc;
	if (!ci || !val) return acc;
#pragma omp parallel for schedule(static) private(cols, vals)
	for (long long r = 0; r < rows; r++)
	{
		const int c = st_row(kind, g, row0 + r, cols, vals);
		double s = 0.0;
		for (int j = 0; j < c; j++)
		{
			ci[rp[r] + j] = (int)cols[j]; val[rp[r] + j] = vals[j];
			s += vals[j] * ((double)(unsigned int)((unsigned long long)cols[j] * 2654435761ull) / 4294967296.0);
		}
		if (b) b[r] = s;
	}
	return acc;
}
