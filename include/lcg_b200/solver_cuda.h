// lcg_b200/solver_cuda.h — C++ drop-in for the class wrappers of liblcg's src/lib/solver_cuda.h:
//   LCG_CUDA_Solver   (solver_cuda.h:35-207,  solver_cuda.cu:29-180)   real double
//   CLCG_CUDA_Solver  (solver_cuda.h:380-541, solver_cuda.cu:300-414)  complex double
// Same usage: derive, implement AxProduct (and MxProduct for the preconditioned call), optionally override Progress, call
// Minimize / MinimizePreconditioned / MinimizeConstrained with HOST vectors.  Header-only; both classes are one template
// over a small traits struct and forward to the entry points of lcg_cuda.h / clcg_cuda.h.
//
// Differences, on purpose: _MxProduct dispatches to MxProduct (the reference's real class calls AxProduct there,
// solver_cuda.h:87-91); wall-clock timing instead of clock(); and one addition — use_builtin_operator(A) makes the
// Minimize* calls run on the fused built-in CSR operator (sentinel callbacks) instead of the virtual AxProduct/MxProduct.
#ifndef LCG_B200_SOLVER_CUDA_H
#define LCG_B200_SOLVER_CUDA_H

#include <chrono>
#include <iostream>
#include "lcg_cuda.h"
#include "clcg_cuda.h"
#include "clcg_cudaf.h"

namespace lcg_b200_detail {

struct RealTraits {
	typedef lcg_float value_t; typedef lcg_para para_t; typedef lcg_solver_enum id_t;
	static para_t defaults() { return lcg_default_parameters(); }
	static void report(int code, bool er_throw) { lcg_error_str(code, er_throw); }
	static const char* name(int id)
	{
		static const char* n[] = {"CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"};
		return (id >= 0 && id < 7) ? n[id] : "Unknown";
	}
};
struct ComplexTraits {
	typedef cuDoubleComplex value_t; typedef clcg_para para_t; typedef clcg_solver_enum id_t;
	static para_t defaults() { return clcg_default_parameters(); }
	static void report(int code, bool er_throw) { clcg_error_str(code, er_throw); }
	static const char* name(int id)
	{
		static const char* n[] = {"BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR", "PCG", "PBICG"};
		return (id >= 0 && id < 7) ? n[id] : "Unknown";
	}
};

template <class Tr>
class SolverBase {
protected:
	typename Tr::para_t param_;
	unsigned int inter_;
	bool silent_;
	lcgb200_csr_t builtin_;

	// prints like the reference's default monitor (solver_cuda.cu:36-51): every inter_ iterations, and at convergence
	int default_progress(const double converge, const double epsilon, const int k)
	{
		if ((inter_ > 0 && k % inter_ == 0) || converge <= epsilon) std::clog << "\rIteration-times: " << k << "\tconvergence: " << converge;
		return 0;
	}
	template <class F> void run(const char* what, bool verbose, bool er_throw, F&& call)
	{
		if (silent_)
		{
			const int ret = call(false);
			if (ret < 0) Tr::report(ret, true);
			return;
		}
		const auto t0 = std::chrono::steady_clock::now();
		const int ret = call(true);
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		if (!er_throw) std::clog << std::endl << "Solver: " << what << ". Time cost: " << ms << " ms" << std::endl;
		if (verbose || ret < 0) Tr::report(ret, er_throw);
	}

public:
	SolverBase() : param_(Tr::defaults()), inter_(1), silent_(false), builtin_(nullptr) {}
	virtual ~SolverBase() {}
	void silent() { silent_ = true; }
	void set_report_interval(unsigned int inter) { inter_ = inter; }
	// run the Minimize* calls on the fused built-in operator (created by the caller, see lcgb200_csr_create); nullptr switches back
	void use_builtin_operator(lcgb200_csr_t A) { builtin_ = A; if (A) lcgb200_csr_set_user(A, this); }
};

}  // namespace lcg_b200_detail

class LCG_CUDA_Solver : public lcg_b200_detail::SolverBase<lcg_b200_detail::RealTraits> {
public:
	virtual void AxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax,
		const int n_size, const int nz_size) = 0;
	virtual void MxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Mx,
		const int n_size, const int nz_size) = 0;
	virtual int Progress(const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int nz_size, const int k)
	{
		(void)m; (void)n_size; (void)nz_size;
		return default_progress(converge, param->epsilon, k);
	}

	static void _AxProduct(void* instance, cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int n, const int nz)
	{
		static_cast<LCG_CUDA_Solver*>(instance)->AxProduct(cub, cus, x, Ax, n, nz);
	}
	static void _MxProduct(void* instance, cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int nz)
	{
		static_cast<LCG_CUDA_Solver*>(instance)->MxProduct(cub, cus, x, Mx, n, nz);
	}
	static int _Progress(void* instance, const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n, const int nz, const int k)
	{
		return static_cast<LCG_CUDA_Solver*>(instance)->Progress(m, converge, param, n, nz, k);
	}

	void set_lcg_parameter(const lcg_para& in_param) { param_ = in_param; }

	void Minimize(cublasHandle_t cub, cusparseHandle_t cus, lcg_float* x, lcg_float* b, const int n_size, const int nz_size,
		lcg_solver_enum solver_id = LCG_CG, bool verbose = true, bool er_throw = false)
	{
		run(lcg_b200_detail::RealTraits::name(solver_id), verbose, er_throw, [&](bool monitor) {
			return lcg_solver_cuda(builtin_ ? lcgb200_csr_ax_typed() : &_AxProduct, monitor ? &_Progress : nullptr, x, b, n_size, nz_size, &param_,
				builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}
	void MinimizePreconditioned(cublasHandle_t cub, cusparseHandle_t cus, lcg_float* x, lcg_float* b, const int n_size, const int nz_size,
		lcg_solver_enum solver_id = LCG_PCG, bool verbose = true, bool er_throw = false)
	{
		run("PCG", verbose, er_throw, [&](bool monitor) {
			return lcg_solver_preconditioned_cuda(builtin_ ? lcgb200_csr_ax_typed() : &_AxProduct, builtin_ ? lcgb200_jacobi_mx_typed() : &_MxProduct,
				monitor ? &_Progress : nullptr, x, b, n_size, nz_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}
	void MinimizeConstrained(cublasHandle_t cub, cusparseHandle_t cus, lcg_float* x, const lcg_float* b, const lcg_float* low, const lcg_float* hig,
		const int n_size, const int nz_size, lcg_solver_enum solver_id = LCG_PG, bool verbose = true, bool er_throw = false)
	{
		run(lcg_b200_detail::RealTraits::name(solver_id == LCG_SPG ? LCG_SPG : LCG_PG), verbose, er_throw, [&](bool monitor) {
			return lcg_solver_constrained_cuda(builtin_ ? lcgb200_csr_ax_typed() : &_AxProduct, monitor ? &_Progress : nullptr, x, b, low, hig,
				n_size, nz_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}

private:
	static lcg_axfunc_cuda_ptr lcgb200_csr_ax_typed() { return reinterpret_cast<lcg_axfunc_cuda_ptr>(&lcgb200_csr_ax); }
	static lcg_axfunc_cuda_ptr lcgb200_jacobi_mx_typed() { return reinterpret_cast<lcg_axfunc_cuda_ptr>(&lcgb200_jacobi_mx); }
};

class CLCG_CUDA_Solver : public lcg_b200_detail::SolverBase<lcg_b200_detail::ComplexTraits> {
public:
	virtual void AxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual void MxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Mx,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual int Progress(const cuDoubleComplex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int nz_size, const int k)
	{
		(void)m; (void)n_size; (void)nz_size;
		return default_progress(converge, param->epsilon, k);
	}

	static void _AxProduct(void* instance, cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int n, const int nz,
		cusparseOperation_t oper_t)
	{
		static_cast<CLCG_CUDA_Solver*>(instance)->AxProduct(cub, cus, x, Ax, n, nz, oper_t);
	}
	static void _MxProduct(void* instance, cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int nz,
		cusparseOperation_t oper_t)
	{
		static_cast<CLCG_CUDA_Solver*>(instance)->MxProduct(cub, cus, x, Mx, n, nz, oper_t);
	}
	static int _Progress(void* instance, const cuDoubleComplex* m, const lcg_float converge, const clcg_para* param, const int n, const int nz, const int k)
	{
		return static_cast<CLCG_CUDA_Solver*>(instance)->Progress(m, converge, param, n, nz, k);
	}

	void set_clcg_parameter(const clcg_para& in_param) { param_ = in_param; }

	void Minimize(cublasHandle_t cub, cusparseHandle_t cus, cuDoubleComplex* x, cuDoubleComplex* b, const int n_size, const int nz_size,
		clcg_solver_enum solver_id = CLCG_BICG, bool verbose = true, bool er_throw = false)
	{
		run(lcg_b200_detail::ComplexTraits::name(solver_id), verbose, er_throw, [&](bool monitor) {
			return clcg_solver_cuda(builtin_ ? reinterpret_cast<clcg_axfunc_cuda_ptr>(&lcgb200_csr_cax) : &_AxProduct, monitor ? &_Progress : nullptr,
				x, b, n_size, nz_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}
	void MinimizePreconditioned(cublasHandle_t cub, cusparseHandle_t cus, cuDoubleComplex* x, cuDoubleComplex* b, const int n_size, const int nz_size,
		clcg_solver_enum solver_id = CLCG_PCG, bool verbose = true, bool er_throw = false)
	{
		run(lcg_b200_detail::ComplexTraits::name(solver_id), verbose, er_throw, [&](bool monitor) {
			return clcg_solver_preconditioned_cuda(builtin_ ? reinterpret_cast<clcg_axfunc_cuda_ptr>(&lcgb200_csr_cax) : &_AxProduct,
				builtin_ ? reinterpret_cast<clcg_axfunc_cuda_ptr>(&lcgb200_jacobi_cmx) : &_MxProduct, monitor ? &_Progress : nullptr,
				x, b, n_size, nz_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}
};

// CLCG_CUDAF_Solver (solver_cuda.h:213-374, solver_cuda.cu:182-298): the same wrapper on cuComplex vectors
class CLCG_CUDAF_Solver : public lcg_b200_detail::SolverBase<lcg_b200_detail::ComplexTraits> {
public:
	virtual void AxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual void MxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Mx,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual int Progress(const cuComplex* m, const float converge, const clcg_para* param, const int n_size, const int nz_size, const int k)
	{
		(void)m; (void)n_size; (void)nz_size;
		return default_progress(converge, param->epsilon, k);
	}

	static void _AxProduct(void* instance, cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Ax, const int n, const int nz,
		cusparseOperation_t oper_t)
	{
		static_cast<CLCG_CUDAF_Solver*>(instance)->AxProduct(cub, cus, x, Ax, n, nz, oper_t);
	}
	static void _MxProduct(void* instance, cublasHandle_t cub, cusparseHandle_t cus, cusparseDnVecDescr_t x, cusparseDnVecDescr_t Mx, const int n, const int nz,
		cusparseOperation_t oper_t)
	{
		static_cast<CLCG_CUDAF_Solver*>(instance)->MxProduct(cub, cus, x, Mx, n, nz, oper_t);
	}
	static int _Progress(void* instance, const cuComplex* m, const float converge, const clcg_para* param, const int n, const int nz, const int k)
	{
		return static_cast<CLCG_CUDAF_Solver*>(instance)->Progress(m, converge, param, n, nz, k);
	}

	void set_clcg_parameter(const clcg_para& in_param) { param_ = in_param; }

	void Minimize(cublasHandle_t cub, cusparseHandle_t cus, cuComplex* x, cuComplex* b, const int n_size, const int nz_size,
		clcg_solver_enum solver_id = CLCG_BICG, bool verbose = true, bool er_throw = false)
	{
		run(lcg_b200_detail::ComplexTraits::name(solver_id), verbose, er_throw, [&](bool monitor) {
			return clcg_solver_cuda(builtin_ ? reinterpret_cast<clcg_axfunc_cudaf_ptr>(&lcgb200_csr_cax) : &_AxProduct, monitor ? &_Progress : nullptr,
				x, b, n_size, nz_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}
	void MinimizePreconditioned(cublasHandle_t cub, cusparseHandle_t cus, cuComplex* x, cuComplex* b, const int n_size, const int nz_size,
		clcg_solver_enum solver_id = CLCG_PCG, bool verbose = true, bool er_throw = false)
	{
		run(lcg_b200_detail::ComplexTraits::name(solver_id), verbose, er_throw, [&](bool monitor) {
			return clcg_solver_preconditioned_cuda(builtin_ ? reinterpret_cast<clcg_axfunc_cudaf_ptr>(&lcgb200_csr_cax) : &_AxProduct,
				builtin_ ? reinterpret_cast<clcg_axfunc_cudaf_ptr>(&lcgb200_jacobi_cmx) : &_MxProduct, monitor ? &_Progress : nullptr,
				x, b, n_size, nz_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, cub, cus, solver_id);
		});
	}
};

#endif  // LCG_B200_SOLVER_CUDA_H
