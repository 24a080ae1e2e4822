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
#include <complex>
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

// ---- complex (clcg) class wrappers: solver_cuda.h:380-541 and solver.h:182-283 ------------------------------------------------
struct CUserSystem { int n = 0, nnz = 0; cusparseSpMatDescr_t A = nullptr; void* buf = nullptr; size_t cap = 0; int calls[3] = {0, 0, 0}; };

class CUserSolver : public CLCG_CUDA_Solver
{
public:
	CUserSystem* sys = nullptr;
	void AxProduct(cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int, const int, cusparseOperation_t op) override
	{
		const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0), zero = make_cuDoubleComplex(0.0, 0.0);
		size_t need = 0;
		cusparseSpMV_bufferSize(cus, op, &one, sys->A, x, &zero, Ax, CUDA_C_64F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
		if (need > sys->cap) { cudaFree(sys->buf); cudaMalloc(&sys->buf, need); sys->cap = need; }
		cusparseSpMV(cus, op, &one, sys->A, x, &zero, Ax, CUDA_C_64F, CUSPARSE_SPMV_ALG_DEFAULT, sys->buf);
		sys->calls[(int)op]++;
	}
	void MxProduct(cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t, cusparseDnVecDescr_t, const int, const int, cusparseOperation_t) override {}
};

class CHostSolver : public CLCG_Solver
{
public:
	const std::vector<int>* rp = nullptr; const std::vector<int>* ci = nullptr; const std::vector<lcg_complex>* va = nullptr;
	int calls = 0;
	void AxProduct(const lcg_complex* x, lcg_complex* y, const int n, lcg_matrix_e layout, clcg_complex_e conj) override
	{
		for (int i = 0; i < n; i++) y[i] = lcg_complex(0.0, 0.0);
		for (int i = 0; i < n; i++)
			for (int k = (*rp)[(size_t)i]; k < (*rp)[(size_t)i + 1]; k++)
			{
				const lcg_complex a = conj == Conjugate ? std::conj((*va)[(size_t)k]) : (*va)[(size_t)k];
				if (layout == MatNormal) y[i] += a * x[(*ci)[(size_t)k]]; else y[(*ci)[(size_t)k]] += a * x[i];
			}
		calls++;
	}
};

static int complex_wrappers(const char* path_a, const char* path_b, cublasHandle_t cub, cusparseHandle_t cus)
{
	FILE* fa = std::fopen(path_a, "rb"); FILE* fb = std::fopen(path_b, "rb");
	if (!fa || !fb) { std::fprintf(stderr, "cannot open the complex fixture\n"); return 1; }
	int n = 0, nz = 0, nb = 0;
	if (std::fread(&n, 4, 1, fa) != 1 || std::fread(&nz, 4, 1, fa) != 1) return 1;
	std::vector<int> r((size_t)nz), c((size_t)nz); std::vector<lcg_complex> v((size_t)nz), b((size_t)n), ans((size_t)n);
	for (int k = 0; k < nz; k++) if (std::fread(&r[(size_t)k], 4, 1, fa) != 1 || std::fread(&c[(size_t)k], 4, 1, fa) != 1 || std::fread(&v[(size_t)k], 16, 1, fa) != 1) return 1;
	if (std::fread(b.data(), 16, (size_t)n, fa) != (size_t)n) return 1;
	if (std::fread(&nb, 4, 1, fb) != 1 || nb != n || std::fread(ans.data(), 16, (size_t)n, fb) != (size_t)n) return 1;
	std::fclose(fa); std::fclose(fb);
	std::vector<int> rp((size_t)n + 1, 0), ci((size_t)nz); std::vector<lcg_complex> va((size_t)nz);
	for (int k = 0; k < nz; k++) rp[(size_t)r[(size_t)k] + 1]++;
	for (int i = 0; i < n; i++) rp[(size_t)i + 1] += rp[(size_t)i];
	{ std::vector<int> fill(rp.begin(), rp.end() - 1); for (int k = 0; k < nz; k++) { int d = fill[(size_t)r[(size_t)k]]++; ci[(size_t)d] = c[(size_t)k]; va[(size_t)d] = v[(size_t)k]; } }
	auto cerr = [&](const lcg_complex* x) { double s = 0.0; for (int i = 0; i < n; i++) s += std::norm(x[i] - ans[(size_t)i]); return std::sqrt(s) / (double)n; };

	int fails = 0;
	CUserSystem sys; sys.n = n; sys.nnz = nz;
	int *d_rp, *d_ci; cuDoubleComplex* d_v;
	cudaMalloc((void**)&d_rp, sizeof(int) * (n + 1)); cudaMalloc((void**)&d_ci, sizeof(int) * nz); cudaMalloc((void**)&d_v, sizeof(cuDoubleComplex) * nz);
	cudaMemcpy(d_rp, rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice); cudaMemcpy(d_ci, ci.data(), sizeof(int) * nz, cudaMemcpyHostToDevice);
	cudaMemcpy(d_v, va.data(), sizeof(cuDoubleComplex) * nz, cudaMemcpyHostToDevice);
	cusparseCreateCsr(&sys.A, n, n, nz, d_rp, d_ci, d_v, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_C_64F);
	lcgb200_csr_t builtin = nullptr;
	if (lcgb200_csr_create(&builtin, n, nz, rp.data(), ci.data(), va.data(), LCGB200_COMPLEX, LCGB200_HOST, LCGB200_CSR_TRANSPOSE | LCGB200_CSR_JACOBI) != 0) return 1;
	clcg_para cp = clcg_default_parameters();
	cp.abs_diff = 1;   // sample6.cpp:162-196 setting
	CUserSolver slv; slv.sys = &sys; slv.set_clcg_parameter(cp); slv.silent();
	std::vector<cuDoubleComplex> m((size_t)n);
	for (int path = 0; path < 2; path++)
	{
		slv.use_builtin_operator(path == 1 ? builtin : nullptr);
		for (auto& z : m) z = make_cuDoubleComplex(0.0, 0.0);
		sys.calls[0] = sys.calls[1] = sys.calls[2] = 0;
		slv.Minimize(cub, cus, m.data(), reinterpret_cast<cuDoubleComplex*>(b.data()), n, nz, CLCG_BICG);
		const double e = cerr(reinterpret_cast<const lcg_complex*>(m.data()));
		std::printf("class CLCG_CUDA_Solver BICG %-26s A-calls %d A^H-calls %d avg-error %.3e\n", path == 0 ? "virtual AxProduct" : "built-in fused operator",
			sys.calls[0], sys.calls[2], e);
		if (!(e < 1e-4)) fails++;
		if (path == 0 && (sys.calls[0] < 100 || sys.calls[2] < 100)) fails++;   // A d1 and A^H d2 every iteration (clcg.cpp:170,188)
		if (path == 1 && (sys.calls[0] || sys.calls[2])) fails++;
	}
	slv.use_builtin_operator(builtin);
	for (auto& z : m) z = make_cuDoubleComplex(0.0, 0.0);
	slv.MinimizePreconditioned(cub, cus, m.data(), reinterpret_cast<cuDoubleComplex*>(b.data()), n, nz, CLCG_PCG);
	if (!(cerr(reinterpret_cast<const lcg_complex*>(m.data())) < 1e-4)) fails++;
	// host-callback CLCG_Solver: (layout, conjugate) callback on the generic path, then the built-in operator
	CHostSolver hs; hs.rp = &rp; hs.ci = &ci; hs.va = &va; hs.set_clcg_parameter(cp); hs.silent();
	std::vector<lcg_complex> mh((size_t)n);
	for (int path = 0; path < 2; path++)
	{
		hs.use_builtin_operator(path == 1 ? builtin : nullptr);
		for (auto& z : mh) z = lcg_complex(0.0, 0.0);
		hs.calls = 0;
		hs.Minimize(mh.data(), b.data(), n, CLCG_BICG_SYM);
		const double e = cerr(mh.data());
		std::printf("class CLCG_Solver BICG_SYM %-27s host Ax calls %d avg-error %.3e\n", path == 0 ? "host callback" : "built-in fused operator", hs.calls, e);
		if (!(e < 1e-4)) fails++;
		if (path == 1 && hs.calls != 0) fails++;
	}
	// CLCG_CUDAF_Solver (solver_cuda.h:213-374): the same system in single-precision complex, caller's cusparseSpMV (CUDA_C_32F)
	// callback on the generic path, then the built-in operator; float storage reaches the answer to float accuracy
	{
		std::vector<cuComplex> vf((size_t)nz), bf((size_t)n), mf((size_t)n);
		for (int k = 0; k < nz; k++) vf[(size_t)k] = make_cuComplex((float)va[(size_t)k].real(), (float)va[(size_t)k].imag());
		for (int i = 0; i < n; i++) bf[(size_t)i] = make_cuComplex((float)b[(size_t)i].real(), (float)b[(size_t)i].imag());
		cuComplex* d_vf; cudaMalloc((void**)&d_vf, sizeof(cuComplex) * nz);
		cudaMemcpy(d_vf, vf.data(), sizeof(cuComplex) * nz, cudaMemcpyHostToDevice);
		struct FSys { cusparseSpMatDescr_t A = nullptr; void* buf = nullptr; size_t cap = 0; int calls = 0; } fs;
		cusparseCreateCsr(&fs.A, n, n, nz, d_rp, d_ci, d_vf, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_C_32F);
		struct FSolver : CLCG_CUDAF_Solver {
			FSys* s = nullptr;
			void AxProduct(cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int, const int, cusparseOperation_t op) override
			{
				const cuComplex one = make_cuComplex(1.f, 0.f), zero = make_cuComplex(0.f, 0.f);
				size_t need = 0;
				cusparseSpMV_bufferSize(cus, op, &one, s->A, x, &zero, y, CUDA_C_32F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
				if (need > s->cap) { cudaFree(s->buf); cudaMalloc(&s->buf, need); s->cap = need; }
				cusparseSpMV(cus, op, &one, s->A, x, &zero, y, CUDA_C_32F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
				s->calls++;
			}
			void MxProduct(cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t, cusparseDnVecDescr_t, const int, const int, cusparseOperation_t) override {}
		} fslv;
		fslv.s = &fs;
		clcg_para fp = clcg_default_parameters();
		fp.max_iterations = 300;   // float storage stalls near 1e-7 relative: bound the run, judge by the error
		fslv.set_clcg_parameter(fp); fslv.silent();
		lcgb200_csr_t fb = nullptr;
		if (lcgb200_csr_create(&fb, n, nz, rp.data(), ci.data(), vf.data(), LCGB200_COMPLEX_FLOAT, LCGB200_HOST, LCGB200_CSR_TRANSPOSE | LCGB200_CSR_JACOBI) != 0) return fails + 1;
		for (int path = 0; path < 2; path++)
		{
			fslv.use_builtin_operator(path == 1 ? fb : nullptr);
			for (auto& z : mf) z = make_cuComplex(0.f, 0.f);
			fs.calls = 0;
			try { fslv.Minimize(cub, cus, mf.data(), bf.data(), n, nz, CLCG_BICG_SYM); } catch (const std::runtime_error&) {}   // -1019 (max iterations) raises in silent mode
			double s2 = 0.0;
			for (int i = 0; i < n; i++) s2 += std::norm(lcg_complex(mf[(size_t)i].x, mf[(size_t)i].y) - ans[(size_t)i]);
			const double e = std::sqrt(s2) / (double)n;
			std::printf("class CLCG_CUDAF_Solver BICG_SYM %-24s A-calls %d avg-error %.3e\n", path == 0 ? "virtual AxProduct" : "built-in fused operator", fs.calls, e);
			if (!(e < 5e-3)) fails++;   // single-precision storage: ~1e-3 on this system (double: 2e-6)
			if (path == 0 && fs.calls < 50) fails++;
			if (path == 1 && fs.calls != 0) fails++;
		}
		lcgb200_csr_destroy(fb);
		cusparseDestroySpMat(fs.A); cudaFree(fs.buf); cudaFree(d_vf);
	}
	lcgb200_csr_destroy(builtin);
	cusparseDestroySpMat(sys.A); cudaFree(sys.buf); cudaFree(d_rp); cudaFree(d_ci); cudaFree(d_v);
	return fails;
}

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
	// complex class wrappers on data/case_10K_cA (configs[1]): CLCG_CUDA_Solver with the caller's cusparseSpMV honouring oper_t
	// (generic path) and on the built-in operator (BiCG needs the stored transpose), then the host-callback CLCG_Solver
	if (argc >= 5) fails += complex_wrappers(argv[3], argv[4], cub, cus);
	lcgb200_csr_destroy(builtin);
	cusparseDestroySpMat(sys.A); cudaFree(sys.d_rp); cudaFree(sys.d_ci); cudaFree(sys.d_v); cudaFree(sys.d_diag); cudaFree(sys.buf);
	cublasDestroy(cub); cusparseDestroy(cus);
	std::printf(fails ? "dropin_sample: %d FAILURES\n" : "dropin_sample: ok\n", fails);
	return fails ? 1 : 0;
}
