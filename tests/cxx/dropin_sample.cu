// A liblcg user's program, re-pointed at the drop-in headers: what src/sample/sample8.cu and sample10.cu do
// (read data/case_10K_A, build CSR, solve with lcg_solver_cuda / lcg_solver_preconditioned_cuda), once with the
// caller's OWN cusparseSpMV + divide callbacks (generic path) and once with the built-in fused operator
// (sentinel callbacks).  Prints one line per solve; exit code 0 iff every solve converged to the known answer
// and both paths agree on the iteration count.
//
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tests/cxx/dropin_sample.cu \
//        -L liblcg_b200 -llcgb200 -lcusparse -lcublas -Xlinker -rpath=$PWD/liblcg_b200 -o dropin_sample
//   ./dropin_sample tests/golden/data/case_10K_A tests/golden/data/case_10K_B
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "lcg_b200/solver_cuda.h"   // pulls in lcg_cuda.h, clcg_cuda.h, util.h
#include "lcg_b200/solver.h"        // host-callback API: lcg.h, clcg.h, LCG_Solver, CLCG_Solver

struct Coo { int r, c; double v; };

struct UserSystem
{	// what the samples keep in globals / the instance pointer (sample8.cu:80-94)
	int n = 0, nnz = 0;
	int *d_rp = nullptr, *d_ci = nullptr; double *d_v = nullptr, *d_diag = nullptr;
	cusparseSpMatDescr_t A = nullptr;
	void* buf = nullptr; size_t buf_bytes = 0;
	int calls_ax = 0, calls_mx = 0, calls_pf = 0, last_k = -1;
};

static void user_ax(void* instance, cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int, const int)
{
	UserSystem* s = static_cast<UserSystem*>(instance);
	const double one = 1.0, zero = 0.0;
	size_t need = 0;
	cusparseSpMV_bufferSize(cus, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, s->A, x, &zero, Ax, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
	if (need > s->buf_bytes) { cudaFree(s->buf); cudaMalloc(&s->buf, need); s->buf_bytes = need; }
	cusparseSpMV(cus, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, s->A, x, &zero, Ax, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
	s->calls_ax++;
}

__global__ void divide_by_diag(int n, const double* r, const double* d, double* z)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) z[i] = r[i] / d[i];
}

static void user_mx(void* instance, cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int)
{	// Jacobi, as sample10.cu:100-121 does with lcg_vecDvecD_element_wise
	UserSystem* s = static_cast<UserSystem*>(instance);
	double *px = nullptr, *pz = nullptr;
	cusparseDnVecGetValues(x, (void**)&px);
	cusparseDnVecGetValues(Mx, (void**)&pz);
	divide_by_diag<<<(n + 255) / 256, 256>>>(n, px, s->d_diag, pz);
	s->calls_mx++;
}

static int user_progress(void* instance, const lcg_float*, const lcg_float, const lcg_para*, const int, const int, const int k)
{
	UserSystem* s = static_cast<UserSystem*>(instance);
	s->calls_pf++; s->last_k = k;
	return 0;
}

// the README-advertised way to use liblcg: derive from the class wrapper (solver_cuda.h:35-207)
class UserSolver : public LCG_CUDA_Solver
{
public:
	UserSystem* sys = nullptr;
	int monitor_calls = 0;
	void AxProduct(cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int n, const int nz) override
	{
		user_ax(sys, cub, cus, x, Ax, n, nz);
	}
	void MxProduct(cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int nz) override
	{
		user_mx(sys, cub, cus, x, Mx, n, nz);
	}
	int Progress(const lcg_float*, const lcg_float, const lcg_para*, const int, const int, const int) override { monitor_calls++; return 0; }
};

// the host-callback class wrapper (solver.h:32-177): AxProduct / MxProduct work on HOST arrays, as with the reference's CPU library
class HostUserSolver : public LCG_Solver
{
public:
	const std::vector<int>* rp = nullptr; const std::vector<int>* ci = nullptr; const std::vector<double>* va = nullptr; const std::vector<double>* dg = nullptr;
	int ax_calls = 0, mx_calls = 0;
	void AxProduct(const lcg_float* x, lcg_float* y, const int n) override
	{
		for (int i = 0; i < n; i++) { double s = 0.0; for (int k = (*rp)[(size_t)i]; k < (*rp)[(size_t)i + 1]; k++) s += (*va)[(size_t)k] * x[(*ci)[(size_t)k]]; y[i] = s; }
		ax_calls++;
	}
	void MxProduct(const lcg_float* r, lcg_float* z, const int n) override { for (int i = 0; i < n; i++) z[i] = r[i] / (*dg)[(size_t)i]; mx_calls++; }
};

static double avg_error(const std::vector<double>& x, const std::vector<double>& ans)
{	// the samples' metric: sqrt(sum |x - ans|^2) / N (sample8.cu:66-74)
	double s = 0.0;
	for (size_t i = 0; i < x.size(); i++) s += (x[i] - ans[i]) * (x[i] - ans[i]);
	return std::sqrt(s) / (double)x.size();
}

int main(int argc, char** argv)
{
	if (argc < 3) { std::fprintf(stderr, "usage: %s case_A case_B\n", argv[0]); return 2; }
	FILE* fa = std::fopen(argv[1], "rb"); FILE* fb = std::fopen(argv[2], "rb");
	if (!fa || !fb) { std::fprintf(stderr, "cannot open the fixture files\n"); return 2; }
	int n = 0, nz = 0, nb = 0;
	if (std::fread(&n, 4, 1, fa) != 1 || std::fread(&nz, 4, 1, fa) != 1) return 2;
	std::vector<Coo> coo((size_t)nz);
	std::vector<double> b((size_t)n), ans((size_t)n);
	for (auto& e : coo) if (std::fread(&e.r, 4, 1, fa) != 1 || std::fread(&e.c, 4, 1, fa) != 1 || std::fread(&e.v, 8, 1, fa) != 1) return 2;
	if (std::fread(b.data(), 8, (size_t)n, fa) != (size_t)n) return 2;
	if (std::fread(&nb, 4, 1, fb) != 1 || nb != n || std::fread(ans.data(), 8, (size_t)n, fb) != (size_t)n) return 2;
	std::fclose(fa); std::fclose(fb);

	// COO -> CSR by counting (the shipped files are row-sorted)
	std::vector<int> rp((size_t)n + 1, 0), ci((size_t)nz); std::vector<double> va((size_t)nz), diag((size_t)n, 0.0);
	for (const auto& e : coo) rp[(size_t)e.r + 1]++;
	for (int i = 0; i < n; i++) rp[(size_t)i + 1] += rp[(size_t)i];
	{ std::vector<int> fill(rp.begin(), rp.end() - 1); for (const auto& e : coo) { int d = fill[(size_t)e.r]++; ci[(size_t)d] = e.c; va[(size_t)d] = e.v; if (e.r == e.c) diag[(size_t)e.r] = e.v; } }

	cublasHandle_t cub; cusparseHandle_t cus;
	cublasCreate(&cub); cusparseCreate(&cus);
	UserSystem sys; sys.n = n; sys.nnz = nz;
	cudaMalloc((void**)&sys.d_rp, sizeof(int) * (n + 1)); cudaMalloc((void**)&sys.d_ci, sizeof(int) * nz);
	cudaMalloc((void**)&sys.d_v, sizeof(double) * nz); cudaMalloc((void**)&sys.d_diag, sizeof(double) * n);
	cudaMemcpy(sys.d_rp, rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice);
	cudaMemcpy(sys.d_ci, ci.data(), sizeof(int) * nz, cudaMemcpyHostToDevice);
	cudaMemcpy(sys.d_v, va.data(), sizeof(double) * nz, cudaMemcpyHostToDevice);
	cudaMemcpy(sys.d_diag, diag.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
	cusparseCreateCsr(&sys.A, n, n, nz, sys.d_rp, sys.d_ci, sys.d_v, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F);

	lcgb200_csr_t builtin = nullptr;
	if (lcgb200_csr_create(&builtin, n, nz, rp.data(), ci.data(), va.data(), LCGB200_REAL, LCGB200_HOST, LCGB200_CSR_JACOBI) != 0)
	{ std::fprintf(stderr, "lcgb200_csr_create: %s\n", lcgb200_last_error()); return 1; }
	lcgb200_csr_set_user(builtin, &sys);

	lcg_para para = lcg_default_parameters();
	para.epsilon = 1e-10;
	int fails = 0;
	struct Run { const char* name; lcg_solver_enum id; bool pre; } runs[] = {{"CG", LCG_CG, false}, {"CGS", LCG_CGS, false}, {"PCG", LCG_PCG, true}};
	for (const Run& r : runs)
	{
		int iters[2] = {0, 0};
		for (int path = 0; path < 2; path++)
		{
			std::vector<double> m((size_t)n, 0.0);
			sys.calls_ax = sys.calls_mx = sys.calls_pf = 0; sys.last_k = -1;
			int ret;
			if (path == 0)
				ret = r.pre ? lcg_solver_preconditioned_cuda(user_ax, user_mx, user_progress, m.data(), b.data(), n, nz, &para, &sys, cub, cus)
				            : lcg_solver_cuda(user_ax, user_progress, m.data(), b.data(), n, nz, &para, &sys, cub, cus, r.id);
			else
				ret = r.pre ? lcg_solver_preconditioned_cuda(lcgb200_csr_ax, lcgb200_jacobi_mx, user_progress, m.data(), b.data(), n, nz, &para, builtin, cub, cus)
				            : lcg_solver_cuda(lcgb200_csr_ax, user_progress, m.data(), b.data(), n, nz, &para, builtin, cub, cus, r.id);
			const double err = avg_error(m, ans);
			iters[path] = sys.last_k;
			std::printf("%-4s %-26s ret %d iterations %d Ax-callbacks %d Mx-callbacks %d avg-error %.3e\n", r.name,
				path == 0 ? "user cusparseSpMV callback" : "built-in fused operator", ret, sys.last_k, sys.calls_ax, sys.calls_mx, err);
			if (ret != LCG_CONVERGENCE || !(err < 1e-4))   // eps 1e-10 on squared norms leaves ~1e-5 (CG rel-L2 6.6e-5) { fails++; lcg_error_str(ret); }
			if (path == 1 && sys.calls_ax != 0) fails++;   // the sentinel must never be called
		}
		if (std::abs(iters[0] - iters[1]) > 1) fails++;
	}
	// class wrappers: virtual AxProduct/MxProduct (with the _MxProduct fix), then the same object on the built-in operator
	{
		UserSolver slv; slv.sys = &sys;
		slv.set_lcg_parameter(para);
		slv.set_report_interval(0);
		for (int path = 0; path < 2; path++)
		{
			slv.use_builtin_operator(path == 1 ? builtin : nullptr);
			std::vector<double> m((size_t)n, 0.0);
			sys.calls_ax = sys.calls_mx = 0; slv.monitor_calls = 0;
			slv.MinimizePreconditioned(cub, cus, m.data(), b.data(), n, nz, LCG_PCG, false, false);
			const double err = avg_error(m, ans);
			std::printf("class PCG %-26s Ax-callbacks %d Mx-callbacks %d monitor-calls %d avg-error %.3e\n",
				path == 0 ? "virtual Ax/MxProduct" : "built-in fused operator", sys.calls_ax, sys.calls_mx, slv.monitor_calls, err);
			if (!(err < 1e-4) || slv.monitor_calls < 50) fails++;
			if (path == 0 && (sys.calls_ax == 0 || sys.calls_mx == 0)) fails++;   // MxProduct really is the preconditioner
			if (path == 1 && (sys.calls_ax != 0 || sys.calls_mx != 0)) fails++;
		}
		lcgb200_csr_set_user(builtin, &sys);
		slv.silent();
		std::vector<double> m((size_t)n, 0.0);
		slv.use_builtin_operator(nullptr);
		slv.Minimize(cub, cus, m.data(), b.data(), n, nz, LCG_CGS);
		if (!(avg_error(m, ans) < 1e-3)) fails++;
		bool threw = false;
		lcg_para bad = para; bad.epsilon = -1.0; slv.set_lcg_parameter(bad);
		try { slv.Minimize(cub, cus, m.data(), b.data(), n, nz, LCG_CG); } catch (const std::runtime_error&) { threw = true; }   // silent_ => throws on error (solver_cuda.cu:73-78)
		if (!threw) fails++;
	}
	// host-callback API: LCG_Solver with host Ax/Mx (generic path, staged over PCIe) and on the built-in operator
	{
		HostUserSolver hs; hs.rp = &rp; hs.ci = &ci; hs.va = &va; hs.dg = &diag;
		hs.set_lcg_parameter(para); hs.silent();
		for (int path = 0; path < 2; path++)
		{
			hs.use_builtin_operator(path == 1 ? builtin : nullptr);
			std::vector<double> m((size_t)n, 0.0);
			hs.ax_calls = hs.mx_calls = 0;
			hs.MinimizePreconditioned(m.data(), b.data(), n);
			const double err = avg_error(m, ans);
			std::printf("host-API class PCG %-24s host Ax calls %d Mx calls %d avg-error %.3e\n", path == 0 ? "host callbacks" : "built-in fused operator", hs.ax_calls, hs.mx_calls, err);
			if (!(err < 1e-4)) fails++;
			if (path == 0 && (hs.ax_calls != 100 || hs.mx_calls != 100)) fails++;   // 1 + 99 iterations: exactly the reference's call count, none past convergence
			if (path == 1 && (hs.ax_calls != 0 || hs.mx_calls != 0)) fails++;
		}
		lcgb200_csr_set_user(builtin, &sys);
		std::vector<double> m((size_t)n, 0.0);
		if (lcg_solver(lcgb200_csr_ax_host, nullptr, m.data(), b.data(), n, &para, builtin) != LCG_CONVERGENCE || !(avg_error(m, ans) < 1e-3)) fails++;   // default id: CGS
	}
	// error behaviour mirrors the reference (lcg_cuda.cu:91-98)
	{
		std::vector<double> m((size_t)n, 0.0);
		lcg_para bad = para; bad.epsilon = 2.0;
		if (lcg_solver_cuda(user_ax, nullptr, m.data(), b.data(), n, nz, &bad, &sys, cub, cus) != LCG_INVILAD_EPSILON) fails++;
		if (lcg_solver_cuda(user_ax, nullptr, m.data(), b.data(), n, nz, &para, &sys, nullptr, cus) != LCG_INVALID_POINTER) fails++;
		if (lcg_solver_cuda(user_ax, nullptr, m.data(), b.data(), 0, nz, &para, &sys, cub, cus) != LCG_INVILAD_VARIABLE_SIZE) fails++;
	}
	lcgb200_csr_destroy(builtin);
	cusparseDestroySpMat(sys.A); cudaFree(sys.d_rp); cudaFree(sys.d_ci); cudaFree(sys.d_v); cudaFree(sys.d_diag); cudaFree(sys.buf);
	cublasDestroy(cub); cusparseDestroy(cus);
	std::printf(fails ? "dropin_sample: %d FAILURES\n" : "dropin_sample: ok\n", fails);
	return fails ? 1 : 0;
}
