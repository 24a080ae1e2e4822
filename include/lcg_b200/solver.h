// lcg_b200/solver.h — C++ drop-in for the HOST-callback class wrappers of liblcg's src/lib/solver.h:
//   LCG_Solver   (solver.h:32-177, solver.cpp:29-216)    real     -> lcg_solver / lcg_solver_preconditioned / lcg_solver_constrained
//   CLCG_Solver  (solver.h:182-283, solver.cpp:218-310)  complex  -> clcg_solver
// Derive and implement AxProduct (MxProduct) on HOST arrays exactly as with the reference; or call
// use_builtin_operator(A) and let the Minimize* calls run on the GPU's fused built-in operator.  Header-only.
#ifndef LCG_B200_SOLVER_H
#define LCG_B200_SOLVER_H

#include <chrono>
#include <iostream>
#include "lcg.h"
#include "clcg.h"

namespace lcg_b200_detail {

template <class Para>
class HostSolverBase {
protected:
	Para param_;
	unsigned int inter_;
	bool silent_;
	lcgb200_csr_t builtin_;
	int default_progress(const double converge, const double epsilon, const int k)
	{	// like solver.cpp:40-54: a line every inter_ iterations and at convergence
		if ((inter_ > 0 && k % inter_ == 0) || converge <= epsilon) std::clog << "\rIteration-times: " << k << "\tconvergence: " << converge;
		return 0;
	}
	template <class F, class R> void run(const char* what, bool verbose, bool er_throw, R&& report, F&& call)
	{
		if (silent_) { const int ret = call(false); if (ret < 0) report(ret, true); return; }
		const auto t0 = std::chrono::steady_clock::now();
		const int ret = call(true);
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		if (!er_throw) std::clog << std::endl << "Solver: " << what << ". Time cost: " << ms << " ms" << std::endl;
		if (verbose || ret < 0) report(ret, er_throw);
	}
public:
	explicit HostSolverBase(const Para& p) : param_(p), inter_(1), silent_(false), builtin_(nullptr) {}
	virtual ~HostSolverBase() {}
	void silent() { silent_ = true; }
	void set_report_interval(unsigned int inter) { inter_ = inter; }
	void use_builtin_operator(lcgb200_csr_t A) { builtin_ = A; if (A) lcgb200_csr_set_user(A, this); }
};

}  // namespace lcg_b200_detail

class LCG_Solver : public lcg_b200_detail::HostSolverBase<lcg_para> {
public:
	LCG_Solver() : HostSolverBase<lcg_para>(lcg_default_parameters()) {}
	virtual void AxProduct(const lcg_float* a, lcg_float* b, const int num) = 0;
	virtual void MxProduct(const lcg_float* a, lcg_float* b, const int num) = 0;
	virtual int Progress(const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int k)
	{
		(void)m; (void)n_size;
		return default_progress(converge, param->epsilon, k);
	}
	static void _AxProduct(void* instance, const lcg_float* a, lcg_float* b, const int num) { static_cast<LCG_Solver*>(instance)->AxProduct(a, b, num); }
	static void _MxProduct(void* instance, const lcg_float* a, lcg_float* b, const int num) { static_cast<LCG_Solver*>(instance)->MxProduct(a, b, num); }
	static int _Progress(void* instance, const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int k)
	{
		return static_cast<LCG_Solver*>(instance)->Progress(m, converge, param, n_size, k);
	}
	void set_lcg_parameter(const lcg_para& in_param) { param_ = in_param; }

	void Minimize(lcg_float* m, const lcg_float* b, int x_size, lcg_solver_enum solver_id = LCG_CG, bool verbose = true, bool er_throw = false)
	{
		static const char* names[] = {"CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"};
		run((solver_id >= 0 && solver_id < 7) ? names[solver_id] : "Unknown", verbose, er_throw, lcg_error_str, [&](bool monitor) {
			return lcg_solver(builtin_ ? &lcgb200_csr_ax_host : &_AxProduct, monitor ? &_Progress : nullptr, m, b, x_size, &param_,
				builtin_ ? (void*)builtin_ : (void*)this, solver_id);
		});
	}
	void MinimizePreconditioned(lcg_float* m, const lcg_float* b, int x_size, lcg_solver_enum solver_id = LCG_PCG, bool verbose = true, bool er_throw = false)
	{
		run("PCG", verbose, er_throw, lcg_error_str, [&](bool monitor) {
			return lcg_solver_preconditioned(builtin_ ? &lcgb200_csr_ax_host : &_AxProduct, builtin_ ? &lcgb200_jacobi_mx_host : &_MxProduct,
				monitor ? &_Progress : nullptr, m, b, x_size, &param_, builtin_ ? (void*)builtin_ : (void*)this, solver_id);
		});
	}
	void MinimizeConstrained(lcg_float* m, const lcg_float* b, const lcg_float* low, const lcg_float* hig, int x_size,
		lcg_solver_enum solver_id = LCG_PG, bool verbose = true, bool er_throw = false)
	{
		run(solver_id == LCG_SPG ? "SPG" : "PG", verbose, er_throw, lcg_error_str, [&](bool monitor) {
			return lcg_solver_constrained(builtin_ ? &lcgb200_csr_ax_host : &_AxProduct, monitor ? &_Progress : nullptr, m, b, low, hig, x_size, &param_,
				builtin_ ? (void*)builtin_ : (void*)this, solver_id);
		});
	}
};

class CLCG_Solver : public lcg_b200_detail::HostSolverBase<clcg_para> {
public:
	CLCG_Solver() : HostSolverBase<clcg_para>(clcg_default_parameters()) {}
	virtual void AxProduct(const lcg_complex* x, lcg_complex* prod_Ax, const int x_size, lcg_matrix_e layout, clcg_complex_e conjugate) = 0;
	virtual int Progress(const lcg_complex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int k)
	{
		(void)m; (void)n_size;
		return default_progress(converge, param->epsilon, k);
	}
	static void _AxProduct(void* instance, const lcg_complex* x, lcg_complex* prod_Ax, const int x_size, lcg_matrix_e layout, clcg_complex_e conjugate)
	{
		static_cast<CLCG_Solver*>(instance)->AxProduct(x, prod_Ax, x_size, layout, conjugate);
	}
	static int _Progress(void* instance, const lcg_complex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int k)
	{
		return static_cast<CLCG_Solver*>(instance)->Progress(m, converge, param, n_size, k);
	}
	void set_clcg_parameter(const clcg_para& in_param) { param_ = in_param; }

	void Minimize(lcg_complex* m, const lcg_complex* b, int x_size, clcg_solver_enum solver_id = CLCG_CGS, bool verbose = true, bool er_throw = false)
	{
		static const char* names[] = {"BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR", "PCG", "PBICG"};
		run((solver_id >= 0 && solver_id < 7) ? names[solver_id] : "Unknown", verbose, er_throw, clcg_error_str, [&](bool monitor) {
			return clcg_solver(builtin_ ? reinterpret_cast<clcg_axfunc_ptr>(&lcgb200_csr_cax_host) : &_AxProduct, monitor ? &_Progress : nullptr, m, b, x_size,
				&param_, builtin_ ? (void*)builtin_ : (void*)this, solver_id);
		});
	}
};

#endif  // LCG_B200_SOLVER_H
