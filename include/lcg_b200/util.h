// lcg_b200/util.h — C++ drop-in for the types of liblcg's src/lib/util.h that sit on the solver boundary:
// solver ids, return codes and the parameter blocks, with the reference's names and numeric values
// (util.h:32-64 lcg_solver_enum, :69-90 lcg_return_enum, :95-153 lcg_para/defparam, :187-221 clcg_solver_enum,
// :226-242 clcg_return_enum, :247-278 clcg_para/defparam2) plus the three small host helpers callers use
// (util.h:162,171,181 / :287,296,306).  Header-only; everything forwards to the C ABI in lcgb200.h.
#ifndef LCG_B200_UTIL_H
#define LCG_B200_UTIL_H

#include <cstdio>
#include <stdexcept>
#include <string>
#include "../lcgb200.h"

typedef double lcg_float;   // algebra.h:50

enum lcg_solver_enum { LCG_CG, LCG_PCG, LCG_CGS, LCG_BICGSTAB, LCG_BICGSTAB2, LCG_PG, LCG_SPG };

enum lcg_return_enum {
	LCG_SUCCESS = 0, LCG_CONVERGENCE = 0, LCG_STOP, LCG_ALREADY_OPTIMIZIED,
	LCG_UNKNOWN_ERROR = -1024, LCG_INVILAD_VARIABLE_SIZE, LCG_INVILAD_MAX_ITERATIONS, LCG_INVILAD_EPSILON,
	LCG_INVILAD_RESTART_EPSILON, LCG_REACHED_MAX_ITERATIONS, LCG_NULL_PRECONDITION_MATRIX, LCG_NAN_VALUE,
	LCG_INVALID_POINTER, LCG_INVALID_LAMBDA, LCG_INVALID_SIGMA, LCG_INVALID_BETA, LCG_INVALID_MAXIM, LCG_SIZE_NOT_MATCH
};

// same fields, order and offsets (0/8/16/24/32/40/48/56) as the reference's struct: it is passed through the C ABI as is
typedef lcgb200_para lcg_para;
static const lcg_para defparam = {0, 1e-6, 0, 1e-6, 1.0, 0.95, 0.9, 10};

enum clcg_solver_enum { CLCG_BICG, CLCG_BICG_SYM, CLCG_CGS, CLCG_BICGSTAB, CLCG_TFQMR, CLCG_PCG, CLCG_PBICG };

enum clcg_return_enum {
	CLCG_SUCCESS = 0, CLCG_CONVERGENCE = 0, CLCG_STOP, CLCG_ALREADY_OPTIMIZIED,
	CLCG_UNKNOWN_ERROR = -1024, CLCG_INVILAD_VARIABLE_SIZE, CLCG_INVILAD_MAX_ITERATIONS, CLCG_INVILAD_EPSILON,
	CLCG_REACHED_MAX_ITERATIONS, CLCG_NAN_VALUE, CLCG_INVALID_POINTER, CLCG_SIZE_NOT_MATCH, CLCG_UNKNOWN_SOLVER
};

typedef lcgb200_cpara clcg_para;
static const clcg_para defparam2 = {0, 1e-6, 0};

inline lcg_para lcg_default_parameters() { return defparam; }
inline clcg_para clcg_default_parameters() { return defparam2; }

// name -> id; unknown names fall back to CGS like util.cpp:39-51 (the complex table only knows four names, util.cpp:157-166)
inline lcg_solver_enum lcg_select_solver(const std::string& name)
{
	static const char* names[] = {"LCG_CG", "LCG_PCG", "LCG_CGS", "LCG_BICGSTAB", "LCG_BICGSTAB2", "LCG_PG", "LCG_SPG"};
	for (int i = 0; i < 7; i++) if (name == names[i]) return static_cast<lcg_solver_enum>(i);
	return LCG_CGS;
}
inline clcg_solver_enum clcg_select_solver(const std::string& name)
{
	if (name == "CLCG_BICG") return CLCG_BICG;
	if (name == "CLCG_BICG_SYM") return CLCG_BICG_SYM;
	if (name == "CLCG_TFQMR") return CLCG_TFQMR;
	return CLCG_CGS;
}

// one line per return code on stderr; with er_throw a negative code raises std::runtime_error (util.cpp:53-148)
inline const char* lcg_b200_code_text(int code, bool cplx)
{
	if (code == 0) return "The iteration reached convergence.";
	if (code == 1) return "The iteration was stopped by the progress callback.";
	if (code == 2) return "The initial solution is already optimized.";
	if (cplx)
	{
		switch (code)
		{
			case CLCG_INVILAD_VARIABLE_SIZE: return "The variable size is not positive.";
			case CLCG_INVILAD_MAX_ITERATIONS: return "The maximal iteration count is negative.";
			case CLCG_INVILAD_EPSILON: return "The epsilon is not in (0,1).";
			case CLCG_REACHED_MAX_ITERATIONS: return "The iteration reached the maximal limit.";
			case CLCG_NAN_VALUE: return "The model values are NaN.";   // also what a max-iteration exit prints: the complex solvers return -1019
			case CLCG_INVALID_POINTER: return "Invalid pointer.";
			case CLCG_SIZE_NOT_MATCH: return "The sizes of the operator and the vectors do not match.";
			case CLCG_UNKNOWN_SOLVER: return "Unknown solver.";
			default: return "Unknown error.";
		}
	}
	switch (code)
	{
		case LCG_INVILAD_VARIABLE_SIZE: return "The variable size is not positive.";
		case LCG_INVILAD_MAX_ITERATIONS: return "The maximal iteration count is negative.";
		case LCG_INVILAD_EPSILON: return "The epsilon is not in (0,1).";
		case LCG_INVILAD_RESTART_EPSILON: return "The restart epsilon is not positive.";
		case LCG_REACHED_MAX_ITERATIONS: return "The iteration reached the maximal limit.";
		case LCG_NULL_PRECONDITION_MATRIX: return "The preconditioner is missing.";
		case LCG_NAN_VALUE: return "The model values are NaN.";
		case LCG_INVALID_POINTER: return "Invalid pointer.";
		case LCG_INVALID_LAMBDA: return "Invalid range for lambda (step).";
		case LCG_INVALID_SIGMA: return "Invalid range for sigma.";
		case LCG_INVALID_BETA: return "Invalid range for beta.";
		case LCG_INVALID_MAXIM: return "Invalid range for maxi_m.";
		case LCG_SIZE_NOT_MATCH: return "The sizes of the operator and the vectors do not match.";
		default: return "Unknown error.";
	}
}
inline void lcg_error_str(int er_index, bool er_throw = false)
{
	const char* text = lcg_b200_code_text(er_index, false);
	if (er_throw && er_index < 0) throw std::runtime_error(std::string("[LibLCG] ") + text);
	std::fprintf(stderr, "%s %s\n", er_index >= 0 ? "Success!" : (er_index == LCG_REACHED_MAX_ITERATIONS ? "Warning!" : "Fail!"), text);
}
inline void clcg_error_str(int er_index, bool er_throw = false)
{
	const char* text = lcg_b200_code_text(er_index, true);
	if (er_throw && er_index < 0) throw std::runtime_error(std::string("[LibLCG] ") + text);
	std::fprintf(stderr, "%s %s\n", er_index >= 0 ? "Success!" : "Fail!", text);
}

#endif  // LCG_B200_UTIL_H
