// A liblcg user's program compiled against the REFERENCE'S OWN HEADERS (-I <liblcg>/src/lib, -DLibLCG_CUDA: lcg.h, clcg.h,
// lcg_cuda.h, clcg_cuda.h, solver.h, solver_cuda.h, util.h, algebra.h, lcg_complex.h — nothing of ours but the C header
// that declares the sentinel callbacks) and linked against liblcg_dropin.so instead of liblcg.so.  Every reference symbol
// it uses must resolve in our library and behave as the reference documents it:
//   lcg_malloc / lcg_vecset / lcg_free / lcg_dot, lcg_default_parameters, lcg_error_str          (algebra.h, util.h)
//   lcg_solver, lcg_solver_preconditioned, lcg(), lcgs() with caller-owned work vectors           (lcg.h:71-169)
//   lcg_solver_cuda with the user's cusparseSpMV callback; lcg_solver_preconditioned_cuda with the built-in operator
//   LCG_Solver / CLCG_Solver / LCG_CUDA_Solver / CLCG_CUDA_Solver subclasses (vtable layout of solver.h / solver_cuda.h)
//   clcg_solver, clcg_solver_cuda, clcg_solver_preconditioned_cuda on data/case_1K_cA
// Built by tests/cxx/Makefile only where the reference tree exists; the binary travels to the GPU box.
//   ./ref_header_sample case_10K_A case_10K_B case_1K_cA case_1K_cB
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "lcg.h"
#include "clcg.h"
#include "lcg_cuda.h"
#include "clcg_cuda.h"
#include "solver.h"
#include "solver_cuda.h"
#include "preconditioner.h"
#include "algebra_cuda.h"
#include "lcg_complex_cuda.h"
#include "lcgb200.h"

struct Csr { int n = 0, nnz = 0; std::vector<int> rp, ci; std::vector<double> va, diag, b, ans; };
struct CCsr { int n = 0, nnz = 0; std::vector<int> rp, ci; std::vector<lcg_complex> va, diag, b, ans; };

template <class V, class M>
static bool load_case(const char* fa_path, const char* fb_path, M& A)
{	// data/README:1-10: int N, int nz, nz x {int row, int col, value}, N x value; the *_B file: int N, N x value
	FILE* fa = std::fopen(fa_path, "rb"); FILE* fb = std::fopen(fb_path, "rb");
	if (!fa || !fb) return false;
	int nb = 0;
	if (std::fread(&A.n, 4, 1, fa) != 1 || std::fread(&A.nnz, 4, 1, fa) != 1) return false;
	std::vector<int> r((size_t)A.nnz), c((size_t)A.nnz); std::vector<V> v((size_t)A.nnz);
	for (int k = 0; k < A.nnz; k++)
		if (std::fread(&r[(size_t)k], 4, 1, fa) != 1 || std::fread(&c[(size_t)k], 4, 1, fa) != 1 || std::fread(&v[(size_t)k], sizeof(V), 1, fa) != 1) return false;
	A.b.resize((size_t)A.n); A.ans.resize((size_t)A.n);
	if (std::fread(A.b.data(), sizeof(V), (size_t)A.n, fa) != (size_t)A.n) return false;
	if (std::fread(&nb, 4, 1, fb) != 1 || nb != A.n || std::fread(A.ans.data(), sizeof(V), (size_t)A.n, fb) != (size_t)A.n) return false;
	std::fclose(fa); std::fclose(fb);
	A.rp.assign((size_t)A.n + 1, 0); A.ci.resize((size_t)A.nnz); A.va.resize((size_t)A.nnz); A.diag.assign((size_t)A.n, V(0));
	for (int k = 0; k < A.nnz; k++) A.rp[(size_t)r[(size_t)k] + 1]++;
	for (int i = 0; i < A.n; i++) A.rp[(size_t)i + 1] += A.rp[(size_t)i];
	std::vector<int> fill(A.rp.begin(), A.rp.end() - 1);
	for (int k = 0; k < A.nnz; k++)
	{
		const int d = fill[(size_t)r[(size_t)k]]++;
		A.ci[(size_t)d] = c[(size_t)k]; A.va[(size_t)d] = v[(size_t)k];
		if (r[(size_t)k] == c[(size_t)k]) A.diag[(size_t)r[(size_t)k]] = v[(size_t)k];
	}
	return true;
}

static Csr g_A; static CCsr g_C;
static int g_ax_calls = 0, g_last_k = -1;

static void host_ax(void*, const lcg_float* x, lcg_float* y, const int n)
{
	for (int i = 0; i < n; i++) { double s = 0.0; for (int k = g_A.rp[(size_t)i]; k < g_A.rp[(size_t)i + 1]; k++) s += g_A.va[(size_t)k] * x[g_A.ci[(size_t)k]]; y[i] = s; }
	g_ax_calls++;
}
static void host_mx(void*, const lcg_float* r, lcg_float* z, const int n) { for (int i = 0; i < n; i++) z[i] = r[i] / g_A.diag[(size_t)i]; }
static int host_pf(void*, const lcg_float*, const lcg_float, const lcg_para*, const int, const int k) { g_last_k = k; return 0; }

// IC(0) preconditioner from the reference's preconditioner.h functions: factor once, two COO triangular solves per application
static std::vector<int> g_lrow, g_lcol, g_urow, g_ucol;
static std::vector<double> g_lval, g_uval, g_tmp;
static int g_lnz = 0;
static void host_ic_mx(void*, const lcg_float* r, lcg_float* z, const int n)
{
	lcg_solve_lower_triangle_coo(g_lrow.data(), g_lcol.data(), g_lval.data(), r, g_tmp.data(), n, g_lnz);
	lcg_solve_upper_triangle_coo(g_urow.data(), g_ucol.data(), g_uval.data(), g_tmp.data(), z, n, g_lnz);
}

static void host_cax(void*, const lcg_complex* x, lcg_complex* y, const int n, lcg_matrix_e layout, clcg_complex_e conj)
{	// op(A) x, honouring (layout, conjugate) as clcg.h:40-41 asks
	for (int i = 0; i < n; i++) y[i] = lcg_complex(0.0, 0.0);
	for (int i = 0; i < n; i++)
		for (int k = g_C.rp[(size_t)i]; k < g_C.rp[(size_t)i + 1]; k++)
		{
			const lcg_complex a = conj == Conjugate ? std::conj(g_C.va[(size_t)k]) : g_C.va[(size_t)k];
			if (layout == MatNormal) y[i] += a * x[g_C.ci[(size_t)k]]; else y[g_C.ci[(size_t)k]] += a * x[i];
		}
}

struct DevSys { cusparseSpMatDescr_t A = nullptr; void* buf = nullptr; size_t cap = 0; double* d_diag = nullptr; int calls = 0; };
static void dev_ax(void* inst, cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int, const int)
{
	DevSys* s = static_cast<DevSys*>(inst);
	const double one = 1.0, zero = 0.0; size_t need = 0;
	cusparseSpMV_bufferSize(cus, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, s->A, x, &zero, y, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
	if (need > s->cap) { cudaFree(s->buf); cudaMalloc(&s->buf, need); s->cap = need; }
	cusparseSpMV(cus, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, s->A, x, &zero, y, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
	s->calls++;
}
// Jacobi Mx callback built from the reference's device helpers, as sample10.cu:117,193 does
static void dev_mx(void* inst, cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int)
{
	DevSys* s = static_cast<DevSys*>(inst);
	void *xp = nullptr, *yp = nullptr;
	cusparseDnVecGetValues(x, &xp); cusparseDnVecGetValues(y, &yp);
	lcg_vecDvecD_element_wise(static_cast<const lcg_float*>(xp), s->d_diag, static_cast<lcg_float*>(yp), n);
}
static void dev_cax(void* inst, cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int, const int, cusparseOperation_t op)
{
	DevSys* s = static_cast<DevSys*>(inst);
	const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0), zero = make_cuDoubleComplex(0.0, 0.0); size_t need = 0;
	cusparseSpMV_bufferSize(cus, op, &one, s->A, x, &zero, y, CUDA_C_64F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
	if (need > s->cap) { cudaFree(s->buf); cudaMalloc(&s->buf, need); s->cap = need; }
	cusparseSpMV(cus, op, &one, s->A, x, &zero, y, CUDA_C_64F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
	s->calls++;
}

// subclasses of the reference's wrapper classes: the object layout and vtable come from the REFERENCE'S headers, the
// out-of-line members (constructor, Minimize*, Progress, ...) from liblcg_dropin.so
class HostSolver : public LCG_Solver {
public:
	void AxProduct(const lcg_float* a, lcg_float* b, const int num) { host_ax(nullptr, a, b, num); }
	void MxProduct(const lcg_float* a, lcg_float* b, const int num) { host_mx(nullptr, a, b, num); }
};
class HostCSolver : public CLCG_Solver {
public:
	void AxProduct(const lcg_complex* x, lcg_complex* y, const int n, lcg_matrix_e l, clcg_complex_e c) { host_cax(nullptr, x, y, n, l, c); }
};
class DevSolver : public LCG_CUDA_Solver {
public:
	DevSys* sys = nullptr;
	void AxProduct(cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz) { dev_ax(sys, cb, cs, x, y, n, nz); }
	void MxProduct(cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t, cusparseDnVecDescr_t, const int, const int) {}
};
class DevCSolver : public CLCG_CUDA_Solver {
public:
	DevSys* sys = nullptr;
	void AxProduct(cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz, cusparseOperation_t op) { dev_cax(sys, cb, cs, x, y, n, nz, op); }
	void MxProduct(cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t, cusparseDnVecDescr_t, const int, const int, cusparseOperation_t) {}
};

static double avg_err(const lcg_float* x, const std::vector<double>& ans) { double s = 0.0; for (size_t i = 0; i < ans.size(); i++) s += (x[i] - ans[i]) * (x[i] - ans[i]); return std::sqrt(s) / (double)ans.size(); }
static double avg_cerr(const lcg_complex* x, const std::vector<lcg_complex>& ans) { double s = 0.0; for (size_t i = 0; i < ans.size(); i++) s += std::norm(x[i] - ans[i]); return std::sqrt(s) / (double)ans.size(); }

int main(int argc, char** argv)
{
	if (argc < 5) { std::fprintf(stderr, "usage: %s case_10K_A case_10K_B case_1K_cA case_1K_cB\n", argv[0]); return 2; }
	if (!load_case<double>(argv[1], argv[2], g_A) || !load_case<lcg_complex>(argv[3], argv[4], g_C)) { std::fprintf(stderr, "cannot read the fixtures\n"); return 2; }
	const int n = g_A.n, nz = g_A.nnz;
	int fails = 0;
	auto check = [&](bool ok, const char* what) { std::printf("%s %s\n", ok ? "ok  " : "FAIL", what); if (!ok) fails++; };

	lcg_para para = lcg_default_parameters();
	check(para.epsilon == 1e-6 && para.maxi_m == 10 && para.sigma == 0.95, "lcg_default_parameters (util.h:153)");
	para.epsilon = 1e-10;
	lcg_float* m = lcg_malloc(n);

	// ---- host-callback API
	lcg_vecset(m, 0.0, n); g_ax_calls = 0;
	int ret = lcg_solver(host_ax, host_pf, m, g_A.b.data(), n, &para, nullptr, LCG_CG);
	check(ret == LCG_CONVERGENCE && g_last_k == 100 && g_ax_calls == 101 && avg_err(m, g_A.ans) < 1e-4, "lcg_solver LCG_CG, host Ax callback: 100 iterations, 1 + 100 Ax calls");
	lcg_vecset(m, 0.0, n);
	ret = lcg_solver_preconditioned(host_ax, host_mx, host_pf, m, g_A.b.data(), n, &para, nullptr);
	check(ret == LCG_CONVERGENCE && g_last_k == 99 && avg_err(m, g_A.ans) < 1e-4, "lcg_solver_preconditioned, host Ax + Jacobi Mx: 99 iterations");
	{	// lcg() with caller-owned work vectors: Gk must come back as the gradient A m - B of the returned solution
		lcg_float *Gk = lcg_malloc(n), *Dk = lcg_malloc(n), *ADk = lcg_malloc(n), *Am = lcg_malloc(n);
		lcg_vecset(m, 0.0, n); lcg_vecset(Gk, 7.0, n);
		ret = lcg(host_ax, nullptr, m, g_A.b.data(), n, &para, nullptr, Gk, Dk, ADk);
		host_ax(nullptr, m, Am, n);
		double dg = 0.0, ng = 0.0;
		for (int i = 0; i < n; i++) { const double g = Am[i] - g_A.b[(size_t)i]; dg += (g - Gk[i]) * (g - Gk[i]); ng += g * g; }
		check(ret == LCG_CONVERGENCE && avg_err(m, g_A.ans) < 1e-4 && std::sqrt(dg) <= 1e-6 * std::sqrt(ng) + 1e-12, "lcg() stand-alone CG, Gk/Dk/ADk owned by the caller (lcg.h:135-137)");
		lcg_float *ws[7]; for (auto& w : ws) w = lcg_malloc(n);
		lcg_vecset(m, 0.0, n);
		ret = lcgs(host_ax, nullptr, m, g_A.b.data(), n, &para, nullptr, ws[0], ws[1], ws[2], ws[3], ws[4], ws[5], ws[6]);
		host_ax(nullptr, m, Am, n);
		lcg_float dr = 0.0, nb = 0.0;   // RK must come back as the residual B - A m of the returned solution
		for (int i = 0; i < n; i++) { const double t = g_A.b[(size_t)i] - Am[i] - ws[0][i]; dr += t * t; }
		lcg_dot(nb, g_A.b.data(), g_A.b.data(), n);
		check(ret == LCG_CONVERGENCE && avg_err(m, g_A.ans) < 1e-3 && std::sqrt(dr) <= 1e-8 * std::sqrt(nb), "lcgs() stand-alone CGS with 7 caller-owned work vectors (lcg.h:166-169); RK holds the final residual");
		for (auto& w : ws) lcg_free(w);
		lcg_free(Gk); lcg_free(Dk); lcg_free(ADk); lcg_free(Am);
	}
	{	// preconditioner.h: IC(0) on the row-sorted COO of case_10K_A, applied through the COO triangular solves inside the reference's PCG
		std::vector<int> row((size_t)nz);
		for (int i = 0; i < n; i++) for (int k = g_A.rp[(size_t)i]; k < g_A.rp[(size_t)i + 1]; k++) row[(size_t)k] = i;
		lcg_incomplete_Cholesky_half_buffsize_coo(row.data(), g_A.ci.data(), nz, &g_lnz);
		g_lrow.resize((size_t)g_lnz); g_lcol.resize((size_t)g_lnz); g_lval.resize((size_t)g_lnz); g_tmp.resize((size_t)n);
		lcg_incomplete_Cholesky_half_coo(row.data(), g_A.ci.data(), g_A.va.data(), n, nz, g_lnz, g_lrow.data(), g_lcol.data(), g_lval.data());
		std::vector<int> cnt((size_t)n + 1, 0);
		for (int k = 0; k < g_lnz; k++) cnt[(size_t)g_lcol[(size_t)k] + 1]++;
		for (int i = 0; i < n; i++) cnt[(size_t)i + 1] += cnt[(size_t)i];
		g_urow.resize((size_t)g_lnz); g_ucol.resize((size_t)g_lnz); g_uval.resize((size_t)g_lnz);
		for (int k = 0; k < g_lnz; k++) { const int d = cnt[(size_t)g_lcol[(size_t)k]]++; g_urow[(size_t)d] = g_lcol[(size_t)k]; g_ucol[(size_t)d] = g_lrow[(size_t)k]; g_uval[(size_t)d] = g_lval[(size_t)k]; }
		lcg_vecset(m, 0.0, n);
		ret = lcg_solver_preconditioned(host_ax, host_ic_mx, host_pf, m, g_A.b.data(), n, &para, nullptr);
		check(ret == LCG_CONVERGENCE && g_last_k == 30 && avg_err(m, g_A.ans) < 1e-4 && lcg_full_rank_coo(g_lrow.data(), g_lcol.data(), g_lval.data(), n, g_lnz),
			"preconditioner.h: lcg_incomplete_Cholesky_half_coo + COO triangular solves as Mx: PCG converges in 30 iterations (Jacobi: 99)");
		// the same preconditioner built into the operator: factor on the host at creation, two level-ordered triangular solves on the GPU
		lcgb200_csr_t opi = nullptr;
		lcgb200_csr_create(&opi, n, nz, g_A.rp.data(), g_A.ci.data(), g_A.va.data(), LCGB200_REAL, LCGB200_HOST, LCGB200_CSR_IC0);
		lcg_vecset(m, 0.0, n);
		ret = lcg_solver_preconditioned(lcgb200_csr_ax_host, lcgb200_ic0_mx_host, host_pf, m, g_A.b.data(), n, &para, opi);
		check(ret == LCG_CONVERGENCE && g_last_k == 30 && avg_err(m, g_A.ans) < 1e-4, "built-in IC(0) (lcgb200_ic0_mx_host): same 30 iterations on the GPU");
		lcgb200_csr_destroy(opi);
	}
	{
		HostSolver hs; hs.set_lcg_parameter(para); hs.silent();
		lcg_vecset(m, 0.0, n);
		hs.MinimizePreconditioned(m, g_A.b.data(), n);
		check(avg_err(m, g_A.ans) < 1e-4, "LCG_Solver subclass (solver.h:32-177) MinimizePreconditioned");
		bool threw = false;
		lcg_para bad = para; bad.epsilon = 5.0; hs.set_lcg_parameter(bad);
		try { hs.Minimize(m, g_A.b.data(), n); } catch (const std::exception&) { threw = true; }
		check(threw, "silent solver raises on a bad parameter through lcg_error_str(er_throw) (solver.cpp:85-92)");
	}
	// ---- complex host API on case_1K_cA
	{
		const int nc = g_C.n;
		clcg_para cp = clcg_default_parameters(); cp.abs_diff = 1;
		lcg_complex* mc = clcg_malloc(nc);
		clcg_vecset(mc, lcg_complex(0.0, 0.0), nc);
		ret = clcg_solver(host_cax, nullptr, mc, g_C.b.data(), nc, &cp, nullptr, CLCG_BICG);
		check(ret == CLCG_CONVERGENCE && avg_cerr(mc, g_C.ans) < 1e-3, "clcg_solver CLCG_BICG with a host (layout, conjugate) callback");
		HostCSolver hc; hc.set_clcg_parameter(cp); hc.silent();
		clcg_vecset(mc, lcg_complex(0.0, 0.0), nc);
		hc.Minimize(mc, g_C.b.data(), nc, CLCG_BICG_SYM);
		check(avg_cerr(mc, g_C.ans) < 1e-3, "CLCG_Solver subclass (solver.h:182-283) Minimize CLCG_BICG_SYM");
		clcg_free(mc);
	}
	// ---- CUDA API
	cublasHandle_t cub; cusparseHandle_t cus;
	cublasCreate(&cub); cusparseCreate(&cus);
	{
		DevSys sys; int *d_rp, *d_ci; double* d_v;
		cudaMalloc((void**)&d_rp, sizeof(int) * (n + 1)); cudaMalloc((void**)&d_ci, sizeof(int) * nz); cudaMalloc((void**)&d_v, sizeof(double) * nz);
		cudaMemcpy(d_rp, g_A.rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice); cudaMemcpy(d_ci, g_A.ci.data(), sizeof(int) * nz, cudaMemcpyHostToDevice);
		cudaMemcpy(d_v, g_A.va.data(), sizeof(double) * nz, cudaMemcpyHostToDevice);
		cusparseCreateCsr(&sys.A, n, n, nz, d_rp, d_ci, d_v, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F);
		lcg_vecset(m, 0.0, n);
		ret = lcg_solver_cuda(dev_ax, nullptr, m, g_A.b.data(), n, nz, &para, &sys, cub, cus, LCG_CG);
		check(ret == LCG_CONVERGENCE && sys.calls == 101 && avg_err(m, g_A.ans) < 1e-4, "lcg_solver_cuda with the caller's cusparseSpMV callback: 1 + 100 Ax calls, none after convergence");
		{	// algebra_cuda.h: the diagonal on the device, then z = r / diag in the caller's Mx callback
			cudaMalloc((void**)&sys.d_diag, sizeof(double) * n);
			lcg_smDcsr_get_diagonal(d_rp, d_ci, d_v, n, sys.d_diag);
			std::vector<double> dg((size_t)n);
			cudaMemcpy(dg.data(), sys.d_diag, sizeof(double) * n, cudaMemcpyDeviceToHost);
			bool same = true;
			for (int i = 0; i < n; i++) same = same && dg[(size_t)i] == g_A.diag[(size_t)i];
			lcg_vecset(m, 0.0, n);
			g_last_k = -1;
			ret = lcg_solver_preconditioned_cuda(dev_ax, dev_mx, [](void*, const lcg_float*, const lcg_float, const lcg_para*, const int, const int, const int k) { g_last_k = k; return 0; },
				m, g_A.b.data(), n, nz, &para, &sys, cub, cus);
			check(same && ret == LCG_CONVERGENCE && g_last_k == 99 && avg_err(m, g_A.ans) < 1e-4,
				"lcg_smDcsr_get_diagonal + lcg_vecDvecD_element_wise in the caller's Jacobi Mx callback (algebra_cuda.h; sample10.cu:117,193): 99 iterations");
			// element-wise product, box clamp and the complex helpers
			double *d_a, *d_b; cudaMalloc((void**)&d_a, sizeof(double) * n); cudaMalloc((void**)&d_b, sizeof(double) * n);
			std::vector<double> ha((size_t)n), hb((size_t)n), hc((size_t)n);
			for (int i = 0; i < n; i++) { ha[(size_t)i] = 0.001 * i - 3.0; hb[(size_t)i] = 1.5 + (i % 7); }
			cudaMemcpy(d_a, ha.data(), sizeof(double) * n, cudaMemcpyHostToDevice); cudaMemcpy(d_b, hb.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
			lcg_vecMvecD_element_wise(d_a, d_b, d_a, n);
			cudaMemcpy(hc.data(), d_a, sizeof(double) * n, cudaMemcpyDeviceToHost);
			bool ok = true;
			for (int i = 0; i < n; i++) ok = ok && hc[(size_t)i] == ha[(size_t)i] * hb[(size_t)i];
			std::vector<double> lo((size_t)n, -2.0), hi((size_t)n, 4.0);
			cudaMemcpy(d_b, lo.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
			double* d_hi; cudaMalloc((void**)&d_hi, sizeof(double) * n); cudaMemcpy(d_hi, hi.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
			lcg_set2box_cuda(d_b, d_hi, d_a, n);
			std::vector<double> hd((size_t)n);
			cudaMemcpy(hd.data(), d_a, sizeof(double) * n, cudaMemcpyDeviceToHost);
			for (int i = 0; i < n; i++) ok = ok && hd[(size_t)i] == std::min(std::max(hc[(size_t)i], -2.0), 4.0);
			cuDoubleComplex *d_z, *d_w;
			cudaMalloc((void**)&d_z, sizeof(cuDoubleComplex) * n); cudaMalloc((void**)&d_w, sizeof(cuDoubleComplex) * n);
			std::vector<cuDoubleComplex> hz((size_t)n), hw((size_t)n), hq((size_t)n), hcj((size_t)n);
			for (int i = 0; i < n; i++) { hz[(size_t)i] = make_cuDoubleComplex(1.0 + 0.01 * i, -2.0 + 0.003 * i); hw[(size_t)i] = make_cuDoubleComplex(0.5 + (i % 5), 1.25 - (i % 3)); }
			cudaMemcpy(d_z, hz.data(), sizeof(cuDoubleComplex) * n, cudaMemcpyHostToDevice); cudaMemcpy(d_w, hw.data(), sizeof(cuDoubleComplex) * n, cudaMemcpyHostToDevice);
			clcg_vecDvecZ_element_wise(d_z, d_w, d_w, n);
			cudaMemcpy(hq.data(), d_w, sizeof(cuDoubleComplex) * n, cudaMemcpyDeviceToHost);
			clcg_vecZ_conjugate(d_z, d_z, n);
			cudaMemcpy(hcj.data(), d_z, sizeof(cuDoubleComplex) * n, cudaMemcpyDeviceToHost);
			for (int i = 0; i < n; i++)
			{
				const std::complex<double> q = std::complex<double>(hz[(size_t)i].x, hz[(size_t)i].y) / std::complex<double>(hw[(size_t)i].x, hw[(size_t)i].y);
				ok = ok && std::abs(std::complex<double>(hq[(size_t)i].x, hq[(size_t)i].y) - q) <= 1e-14 * std::abs(q);
				ok = ok && hcj[(size_t)i].x == hz[(size_t)i].x && hcj[(size_t)i].y == -hz[(size_t)i].y;
			}
			// host helpers: transpose of a row-sorted COO matrix (a repeated entry keeps its last value), cuComplex arithmetic
			const int tr[5] = {0, 0, 1, 2, 2}, tc[5] = {0, 2, 1, 0, 0};
			const cuDoubleComplex tv[5] = {{1, 0}, {2, 0}, {3, 0}, {4, 0}, {5, 1}};
			int orow[5] = {-1, -1, -1, -1, -1}, ocol[5] = {-1, -1, -1, -1, -1}; cuDoubleComplex ov[5] = {};
			clcg_smZcoo_row2col(tr, tc, tv, 3, 5, orow, ocol, ov);
			ok = ok && orow[0] == 0 && ocol[0] == 0 && ov[0].x == 1 && orow[1] == 0 && ocol[1] == 2 && ov[1].x == 5 && ov[1].y == 1 && orow[2] == 1 && ocol[2] == 1 &&
				orow[3] == 2 && ocol[3] == 0 && ov[3].x == 2 && orow[4] == -1;
			const cuDoubleComplex zs = clcg_Zsqrt(make_cuDoubleComplex(-4.0, 0.0)), zd = clcg_Zdiff(clcg_Zsum(tv[0], tv[4]), clcg_Zscale(2.0, tv[1]));
			ok = ok && std::fabs(zs.x) < 1e-15 && zs.y == 2.0 && zd.x == 2.0 && zd.y == 1.0 && cuda2lcg_complex(lcg2cuda_complex(lcg_complex(1.5, -2.5))) == lcg_complex(1.5, -2.5);
			check(ok, "lcg_vecMvecD_element_wise, lcg_set2box_cuda, clcg_vecDvecZ_element_wise, clcg_vecZ_conjugate, clcg_smZcoo_row2col, clcg_Z* (algebra_cuda.h, lcg_complex_cuda.h)");
			cudaFree(d_a); cudaFree(d_b); cudaFree(d_hi); cudaFree(d_z); cudaFree(d_w); cudaFree(sys.d_diag); sys.d_diag = nullptr;
		}
		lcgb200_csr_t op = nullptr;
		lcgb200_csr_create(&op, n, nz, g_A.rp.data(), g_A.ci.data(), g_A.va.data(), LCGB200_REAL, LCGB200_HOST, LCGB200_CSR_JACOBI);
		lcg_vecset(m, 0.0, n);
		ret = lcg_solver_preconditioned_cuda(reinterpret_cast<lcg_axfunc_cuda_ptr>(lcgb200_csr_ax), reinterpret_cast<lcg_axfunc_cuda_ptr>(lcgb200_jacobi_mx), nullptr,
			m, g_A.b.data(), n, nz, &para, op, cub, cus);
		check(ret == LCG_CONVERGENCE && avg_err(m, g_A.ans) < 1e-4, "lcg_solver_preconditioned_cuda on the built-in fused operator (sentinel callbacks)");
		lcgb200_csr_destroy(op);
		DevSolver ds; ds.sys = &sys; ds.set_lcg_parameter(para); ds.silent();
		lcg_vecset(m, 0.0, n);
		ds.Minimize(cub, cus, m, g_A.b.data(), n, nz, LCG_CGS);
		check(avg_err(m, g_A.ans) < 1e-3, "LCG_CUDA_Solver subclass (solver_cuda.h:35-207) Minimize LCG_CGS");
		cusparseDestroySpMat(sys.A); cudaFree(sys.buf); cudaFree(d_rp); cudaFree(d_ci); cudaFree(d_v);
	}
	{
		const int nc = g_C.n, nzc = g_C.nnz;
		DevSys sys; int *d_rp, *d_ci; cuDoubleComplex* d_v;
		cudaMalloc((void**)&d_rp, sizeof(int) * (nc + 1)); cudaMalloc((void**)&d_ci, sizeof(int) * nzc); cudaMalloc((void**)&d_v, sizeof(cuDoubleComplex) * nzc);
		cudaMemcpy(d_rp, g_C.rp.data(), sizeof(int) * (nc + 1), cudaMemcpyHostToDevice); cudaMemcpy(d_ci, g_C.ci.data(), sizeof(int) * nzc, cudaMemcpyHostToDevice);
		cudaMemcpy(d_v, g_C.va.data(), sizeof(cuDoubleComplex) * nzc, cudaMemcpyHostToDevice);
		cusparseCreateCsr(&sys.A, nc, nc, nzc, d_rp, d_ci, d_v, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_C_64F);
		clcg_para cp = clcg_default_parameters(); cp.abs_diff = 1;
		std::vector<cuDoubleComplex> mc((size_t)nc, make_cuDoubleComplex(0.0, 0.0));
		ret = clcg_solver_cuda(dev_cax, nullptr, mc.data(), reinterpret_cast<const cuDoubleComplex*>(g_C.b.data()), nc, nzc, &cp, &sys, cub, cus, CLCG_BICG);
		check(ret == CLCG_CONVERGENCE && avg_cerr(reinterpret_cast<const lcg_complex*>(mc.data()), g_C.ans) < 1e-3, "clcg_solver_cuda CLCG_BICG, cusparseSpMV callback honouring oper_t");
		DevCSolver dc; dc.sys = &sys; dc.set_clcg_parameter(cp); dc.silent();
		for (auto& z : mc) z = make_cuDoubleComplex(0.0, 0.0);
		dc.Minimize(cub, cus, mc.data(), reinterpret_cast<cuDoubleComplex*>(g_C.b.data()), nc, nzc, CLCG_BICG_SYM);
		check(avg_cerr(reinterpret_cast<const lcg_complex*>(mc.data()), g_C.ans) < 1e-3, "CLCG_CUDA_Solver subclass (solver_cuda.h:380-541) Minimize CLCG_BICG_SYM");
		lcgb200_csr_t op = nullptr;
		lcgb200_csr_create(&op, nc, nzc, g_C.rp.data(), g_C.ci.data(), g_C.va.data(), LCGB200_COMPLEX, LCGB200_HOST, LCGB200_CSR_JACOBI);
		for (auto& z : mc) z = make_cuDoubleComplex(0.0, 0.0);
		ret = clcg_solver_preconditioned_cuda(reinterpret_cast<clcg_axfunc_cuda_ptr>(lcgb200_csr_cax), reinterpret_cast<clcg_axfunc_cuda_ptr>(lcgb200_jacobi_cmx), nullptr,
			mc.data(), reinterpret_cast<const cuDoubleComplex*>(g_C.b.data()), nc, nzc, &cp, op, cub, cus);
		check(ret == CLCG_CONVERGENCE && avg_cerr(reinterpret_cast<const lcg_complex*>(mc.data()), g_C.ans) < 1e-3, "clcg_solver_preconditioned_cuda (Jacobi) on the built-in operator");
		lcgb200_csr_destroy(op);
		cusparseDestroySpMat(sys.A); cudaFree(sys.buf); cudaFree(d_rp); cudaFree(d_ci); cudaFree(d_v);
	}
	// error codes and lcg_error_str
	check(lcg_solver_cuda(dev_ax, nullptr, m, g_A.b.data(), n, nz, &para, nullptr, nullptr, cus, LCG_CG) == LCG_INVALID_POINTER, "null cuBLAS handle -> LCG_INVALID_POINTER (lcg_cuda.cu:97-98)");
	bool threw = false;
	try { lcg_error_str(LCG_INVILAD_EPSILON, true); } catch (const std::exception&) { threw = true; }
	check(threw, "lcg_error_str(er_throw = true) throws for a negative code (util.cpp:120)");
	lcg_free(m);
	cublasDestroy(cub); cusparseDestroy(cus);
	std::printf(fails ? "ref_header_sample: %d FAILURES\n" : "ref_header_sample: ok\n", fails);
	return fails ? 1 : 0;
}
