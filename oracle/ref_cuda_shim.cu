// TEST / BENCH INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// Driver around the UNMODIFIED reference CUDA solvers (compiled by oracle/Makefile from
// /root/reference/src/lib/{lcg_cuda.cu, algebra_cuda.cu, util.cpp, algebra.cpp} with -DLibLCG_CUDA into
// oracle/_ref/liblcg_ref_cuda.so): the reference's own GPU path — cuBLAS level-1 calls from the host loop plus the
// caller's cusparseSpMV in the Ax callback — run on the same B200 as a "beat that" baseline for bench.py
// (optional third column).  The callbacks do what the reference's samples do by hand:
// cusparseSpMV for Ax (sample8.cu:96-103, with the CUDA-12 generic API) and the element-wise divide by the CSR
// diagonal for the Jacobi Mx (sample10.cu:100-121,193).
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusparse_v2.h>

#include "lcg_cuda.h"       // resolved with -I/root/reference/src/lib at build time (never copied)
#include "algebra_cuda.h"

namespace {

struct Sys
{
	cusparseSpMatDescr_t A = nullptr;
	void* buf = nullptr; size_t buf_bytes = 0;
	double* d_diag = nullptr;
	int last_k = -1;
};

void ref_ax(void* instance, cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int, const int)
{
	Sys* s = static_cast<Sys*>(instance);
	const double one = 1.0, zero = 0.0;
	cusparseSpMV(cus, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, s->A, x, &zero, Ax, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
}

void ref_mx(void* instance, cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int)
{
	Sys* s = static_cast<Sys*>(instance);
	double *px = nullptr, *pz = nullptr;
	cusparseDnVecGetValues(x, (void**)&px);
	cusparseDnVecGetValues(Mx, (void**)&pz);
	lcg_vecDvecD_element_wise(px, s->d_diag, pz, n);   // the reference's own kernel (algebra_cuda.cu:69-77,103-110)
}

int ref_progress(void* instance, const lcg_float*, const lcg_float, const lcg_para*, const int, const int, const int k)
{
	static_cast<Sys*>(instance)->last_k = k;
	return 0;
}

}  // namespace

// CSR arrays are DEVICE pointers; m (in/out) and b are HOST arrays, as the reference API wants them.
// solver: 0 = CG (lcg_solver_cuda), 1 = Jacobi-PCG (lcg_solver_preconditioned_cuda), 2 = CGS.  with_progress: pass a
// progress callback (the reference then records k; its loop is synchronous either way).
// Returns the reference's return code; *seconds = wall time of the solver call, *iterations = last k seen (or -1).
extern "C" int lcgrefcuda_solve(int solver, int n, int nnz, const int* d_rp, const int* d_ci, const double* d_val, double* m, const double* b,
	double epsilon, int max_iterations, int with_progress, double* seconds, int* iterations)
{
	cublasHandle_t cub; cusparseHandle_t cus;
	if (cublasCreate(&cub) != CUBLAS_STATUS_SUCCESS || cusparseCreate(&cus) != CUSPARSE_STATUS_SUCCESS) return -9999;
	Sys sys;
	cusparseCreateCsr(&sys.A, n, n, nnz, const_cast<int*>(d_rp), const_cast<int*>(d_ci), const_cast<double*>(d_val),
		CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F);
	{	// SpMV workspace, sized once (the samples do this before the solve, sample8.cu:179-181)
		double *tx = nullptr, *ty = nullptr;
		cudaMalloc((void**)&tx, sizeof(double) * n); cudaMalloc((void**)&ty, sizeof(double) * n);
		cusparseDnVecDescr_t vx, vy;
		cusparseCreateDnVec(&vx, n, tx, CUDA_R_64F); cusparseCreateDnVec(&vy, n, ty, CUDA_R_64F);
		const double one = 1.0, zero = 0.0;
		cusparseSpMV_bufferSize(cus, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, sys.A, vx, &zero, vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &sys.buf_bytes);
		cudaMalloc(&sys.buf, sys.buf_bytes > 0 ? sys.buf_bytes : 16);
		cusparseDestroyDnVec(vx); cusparseDestroyDnVec(vy); cudaFree(tx); cudaFree(ty);
	}
	if (solver == 1)
	{
		cudaMalloc((void**)&sys.d_diag, sizeof(double) * n);
		lcg_smDcsr_get_diagonal(d_rp, d_ci, d_val, n, sys.d_diag);   // the reference's own kernel (algebra_cuda.cu:40-57)
	}
	lcg_para para = lcg_default_parameters();
	para.epsilon = epsilon; para.max_iterations = max_iterations;
	cudaDeviceSynchronize();
	const auto t0 = std::chrono::steady_clock::now();
	int ret;
	lcg_progress_cuda_ptr pf = with_progress ? ref_progress : nullptr;
	if (solver == 1) ret = lcg_solver_preconditioned_cuda(ref_ax, ref_mx, pf, m, b, n, nnz, &para, &sys, cub, cus);
	else ret = lcg_solver_cuda(ref_ax, pf, m, b, n, nnz, &para, &sys, cub, cus, solver == 2 ? LCG_CGS : LCG_CG);
	cudaDeviceSynchronize();
	if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (iterations) *iterations = sys.last_k;
	cusparseDestroySpMat(sys.A); cudaFree(sys.buf); cudaFree(sys.d_diag);
	cublasDestroy(cub); cusparseDestroy(cus);
	return ret;
}

// ------------------------------------------------------------------------------------------------ complex
// The reference's complex CUDA solvers (clcg_cuda.cu: clbicg :86-252, clbicg_symmetric :254-401, clpcg :403-559),
// unmodified, behind the same kind of driver: cusparseSpMV honouring oper_t for Ax (sample9.cu:96-104) and the element-wise
// divide by the CSR diagonal for the Jacobi Mx (sample10.cu:100-121,193).  Used by tests/ to pin the complex Jacobi-PCG
// and BiCG paths at a fixed number of iterations (the reference has no buildable CPU complex PCG).
#include "clcg_cuda.h"

namespace {

struct CSys
{
	cusparseSpMatDescr_t A = nullptr;
	void* buf = nullptr; size_t buf_bytes = 0;
	cuDoubleComplex* d_diag = nullptr;
	int last_k = -1, calls = 0;
	double* hist = nullptr; int hist_cap = 0;
};

void cref_ax(void* instance, cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int, const int,
	cusparseOperation_t oper_t)
{
	CSys* s = static_cast<CSys*>(instance);
	const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0), zero = make_cuDoubleComplex(0.0, 0.0);
	cusparseSpMV(cus, oper_t, &one, s->A, x, &zero, Ax, CUDA_C_64F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
}

void cref_mx(void* instance, cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int,
	cusparseOperation_t)
{
	CSys* s = static_cast<CSys*>(instance);
	cuDoubleComplex *px = nullptr, *pz = nullptr;
	cusparseDnVecGetValues(x, (void**)&px);
	cusparseDnVecGetValues(Mx, (void**)&pz);
	clcg_vecDvecZ_element_wise(px, s->d_diag, pz, n);   // the reference's own kernel (lcg_complex_cuda.cu)
}

int cref_progress(void* instance, const cuDoubleComplex*, const lcg_float converge, const clcg_para*, const int, const int, const int k)
{
	CSys* s = static_cast<CSys*>(instance);
	s->last_k = k;
	if (s->hist && s->calls < s->hist_cap) s->hist[s->calls] = converge;
	s->calls++;
	return 0;
}

}  // namespace

// solver: 0 = CLCG_BICG, 1 = CLCG_BICG_SYM, 5 = CLCG_PCG (Jacobi).  CSR arrays: DEVICE pointers (values interleaved re,im);
// m (in/out), b: HOST arrays of n complex values.  hist (nullable) receives the `converge` value of every progress call.
extern "C" int lcgrefcuda_csolve(int solver, int n, int nnz, const int* d_rp, const int* d_ci, const void* d_val, void* m, const void* b,
	double epsilon, int max_iterations, int abs_diff, double* hist, int hist_cap, double* seconds, int* iterations, int* calls)
{
	cublasHandle_t cub; cusparseHandle_t cus;
	if (cublasCreate(&cub) != CUBLAS_STATUS_SUCCESS || cusparseCreate(&cus) != CUSPARSE_STATUS_SUCCESS) return -9999;
	CSys sys;
	sys.hist = hist; sys.hist_cap = hist_cap;
	cusparseCreateCsr(&sys.A, n, n, nnz, const_cast<int*>(d_rp), const_cast<int*>(d_ci), const_cast<void*>(d_val),
		CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_C_64F);
	{
		cuDoubleComplex *tx = nullptr, *ty = nullptr;
		cudaMalloc((void**)&tx, sizeof(cuDoubleComplex) * n); cudaMalloc((void**)&ty, sizeof(cuDoubleComplex) * n);
		cusparseDnVecDescr_t vx, vy;
		cusparseCreateDnVec(&vx, n, tx, CUDA_C_64F); cusparseCreateDnVec(&vy, n, ty, CUDA_C_64F);
		const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0), zero = make_cuDoubleComplex(0.0, 0.0);
		const cusparseOperation_t ops[3] = {CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_TRANSPOSE, CUSPARSE_OPERATION_CONJUGATE_TRANSPOSE};
		for (int i = 0; i < 3; i++)
		{
			size_t need = 0;
			cusparseSpMV_bufferSize(cus, ops[i], &one, sys.A, vx, &zero, vy, CUDA_C_64F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
			if (need > sys.buf_bytes) sys.buf_bytes = need;
		}
		cudaMalloc(&sys.buf, sys.buf_bytes > 0 ? sys.buf_bytes : 16);
		cusparseDestroyDnVec(vx); cusparseDestroyDnVec(vy); cudaFree(tx); cudaFree(ty);
	}
	if (solver == 5)
	{
		cudaMalloc((void**)&sys.d_diag, sizeof(cuDoubleComplex) * n);
		clcg_smZcsr_get_diagonal(d_rp, d_ci, static_cast<const cuDoubleComplex*>(d_val), n, sys.d_diag);
	}
	clcg_para para = clcg_default_parameters();
	para.epsilon = epsilon; para.max_iterations = max_iterations; para.abs_diff = abs_diff;
	cudaDeviceSynchronize();
	const auto t0 = std::chrono::steady_clock::now();
	int ret;
	if (solver == 5) ret = clcg_solver_preconditioned_cuda(cref_ax, cref_mx, cref_progress, static_cast<cuDoubleComplex*>(m),
		static_cast<const cuDoubleComplex*>(b), n, nnz, &para, &sys, cub, cus, CLCG_PCG);
	else ret = clcg_solver_cuda(cref_ax, cref_progress, static_cast<cuDoubleComplex*>(m), static_cast<const cuDoubleComplex*>(b), n, nnz, &para, &sys,
		cub, cus, solver == 1 ? CLCG_BICG_SYM : CLCG_BICG);
	cudaDeviceSynchronize();
	if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (iterations) *iterations = sys.last_k;
	if (calls) *calls = sys.calls;
	cusparseDestroySpMat(sys.A); cudaFree(sys.buf); cudaFree(sys.d_diag);
	cublasDestroy(cub); cusparseDestroy(cus);
	return ret;
}

// ------------------------------------------------------------------------------------------ complex, single precision
// The cuComplex overloads (clcg_cudaf.cu: BICG :86-252, BICG_SYM :254-401, PCG :403-558), unmodified, same driver shape.
#include "clcg_cudaf.h"

namespace {

struct FSys
{
	cusparseSpMatDescr_t A = nullptr;
	void* buf = nullptr; size_t buf_bytes = 0;
	cuComplex* d_diag = nullptr;
	int last_k = -1, calls = 0;
	double* hist = nullptr; int hist_cap = 0;
};

void fref_ax(void* instance, cublasHandle_t, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int, const int,
	cusparseOperation_t oper_t)
{
	FSys* s = static_cast<FSys*>(instance);
	const cuComplex one = make_cuComplex(1.f, 0.f), zero = make_cuComplex(0.f, 0.f);
	cusparseSpMV(cus, oper_t, &one, s->A, x, &zero, Ax, CUDA_C_32F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
}

void fref_mx(void* instance, cublasHandle_t, cusparseHandle_t, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int,
	cusparseOperation_t)
{
	FSys* s = static_cast<FSys*>(instance);
	cuComplex *px = nullptr, *pz = nullptr;
	cusparseDnVecGetValues(x, (void**)&px);
	cusparseDnVecGetValues(Mx, (void**)&pz);
	clcg_vecDvecC_element_wise(px, s->d_diag, pz, n);   // the reference's own kernel (lcg_complex_cuda.cu)
}

int fref_progress(void* instance, const cuComplex*, const float converge, const clcg_para*, const int, const int, const int k)
{
	FSys* s = static_cast<FSys*>(instance);
	s->last_k = k;
	if (s->hist && s->calls < s->hist_cap) s->hist[s->calls] = (double)converge;
	s->calls++;
	return 0;
}

}  // namespace

// solver: 0 = CLCG_BICG, 1 = CLCG_BICG_SYM, 5 = CLCG_PCG (Jacobi).  d_val: DEVICE pointer to nnz cuComplex; m (in/out), b: HOST cuComplex.
extern "C" int lcgrefcuda_csolvef(int solver, int n, int nnz, const int* d_rp, const int* d_ci, const void* d_val, void* m, const void* b,
	double epsilon, int max_iterations, int abs_diff, double* hist, int hist_cap, double* seconds, int* iterations, int* calls)
{
	cublasHandle_t cub; cusparseHandle_t cus;
	if (cublasCreate(&cub) != CUBLAS_STATUS_SUCCESS || cusparseCreate(&cus) != CUSPARSE_STATUS_SUCCESS) return -9999;
	FSys sys;
	sys.hist = hist; sys.hist_cap = hist_cap;
	cusparseCreateCsr(&sys.A, n, n, nnz, const_cast<int*>(d_rp), const_cast<int*>(d_ci), const_cast<void*>(d_val),
		CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_C_32F);
	{
		cuComplex *tx = nullptr, *ty = nullptr;
		cudaMalloc((void**)&tx, sizeof(cuComplex) * n); cudaMalloc((void**)&ty, sizeof(cuComplex) * n);
		cusparseDnVecDescr_t vx, vy;
		cusparseCreateDnVec(&vx, n, tx, CUDA_C_32F); cusparseCreateDnVec(&vy, n, ty, CUDA_C_32F);
		const cuComplex one = make_cuComplex(1.f, 0.f), zero = make_cuComplex(0.f, 0.f);
		const cusparseOperation_t ops[3] = {CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_TRANSPOSE, CUSPARSE_OPERATION_CONJUGATE_TRANSPOSE};
		for (int i = 0; i < 3; i++)
		{
			size_t need = 0;
			cusparseSpMV_bufferSize(cus, ops[i], &one, sys.A, vx, &zero, vy, CUDA_C_32F, CUSPARSE_SPMV_ALG_DEFAULT, &need);
			if (need > sys.buf_bytes) sys.buf_bytes = need;
		}
		cudaMalloc(&sys.buf, sys.buf_bytes > 0 ? sys.buf_bytes : 16);
		cusparseDestroyDnVec(vx); cusparseDestroyDnVec(vy); cudaFree(tx); cudaFree(ty);
	}
	if (solver == 5)
	{
		cudaMalloc((void**)&sys.d_diag, sizeof(cuComplex) * n);
		clcg_smCcsr_get_diagonal(d_rp, d_ci, static_cast<const cuComplex*>(d_val), n, sys.d_diag);
	}
	clcg_para para = clcg_default_parameters();
	para.epsilon = epsilon; para.max_iterations = max_iterations; para.abs_diff = abs_diff;
	cudaDeviceSynchronize();
	const auto t0 = std::chrono::steady_clock::now();
	int ret;
	if (solver == 5) ret = clcg_solver_preconditioned_cuda(fref_ax, fref_mx, fref_progress, static_cast<cuComplex*>(m),
		static_cast<const cuComplex*>(b), n, nnz, &para, &sys, cub, cus, CLCG_PCG);
	else ret = clcg_solver_cuda(fref_ax, fref_progress, static_cast<cuComplex*>(m), static_cast<const cuComplex*>(b), n, nnz, &para, &sys,
		cub, cus, solver == 1 ? CLCG_BICG_SYM : CLCG_BICG);
	cudaDeviceSynchronize();
	if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (iterations) *iterations = sys.last_k;
	if (calls) *calls = sys.calls;
	cusparseDestroySpMat(sys.A); cudaFree(sys.buf); cudaFree(sys.d_diag);
	cublasDestroy(cub); cusparseDestroy(cus);
	return ret;
}

// ------------------------------------------------------------------------------------------ complex IC(0), host functions
// clcg_incomplete_Cholesky_cuda_half (preconditioner_cuda.cu:40-270): sequential host code despite its name; used to pin our
// restatement of the complex factorisation bit for bit.  prec: 0 = cuDoubleComplex, 1 = cuComplex.  Returns the size of L.
#include "preconditioner_cuda.h"
extern "C" int lcgrefcuda_cic0_half(int prec, const int* row, const int* col, const void* val, int n, int nz, int* ic_row, int* ic_col, void* ic_val)
{
	int lnz = 0;
	clcg_incomplete_Cholesky_cuda_half_buffsize(row, col, nz, &lnz);
	if (ic_row && ic_col && ic_val)
	{
		if (prec == 0) clcg_incomplete_Cholesky_cuda_half(row, col, static_cast<const cuDoubleComplex*>(val), n, nz, lnz, ic_row, ic_col, static_cast<cuDoubleComplex*>(ic_val));
		else clcg_incomplete_Cholesky_cuda_half(row, col, static_cast<const cuComplex*>(val), n, nz, lnz, ic_row, ic_col, static_cast<cuComplex*>(ic_val));
	}
	return lnz;
}
