// solvers.cuh — entry points of the iteration loops (solvers_real.cu, solvers_complex.cu)
#pragma once
#include "engine.cuh"
#include "../../include/lcgb200.h"

namespace lcgb200 {

// n = local rows, next = length of vectors that are SpMV inputs (n + ghost entries)
int solve_real(Engine& E, const Operator<double>& A, int solver_id, double* m, const double* B, const double* lo, const double* hi,
	const lcgb200_para& para, size_t n, size_t next);
int real_vector_count(int solver_id);

int solve_complex(Engine& E, const Operator<double2>& A, int solver_id, double2* m, const double2* B,
	const lcgb200_cpara& para, size_t n, size_t next);
int complex_vector_count(int solver_id);
// the same loops on cuComplex vectors (clcg_cudaf.h:81-105): storage float2, arithmetic and scalars in double
int solve_complexf(Engine& E, const Operator<ZF>& A, int solver_id, ZF* m, const ZF* B,
	const lcgb200_cpara& para, size_t n, size_t next);

// reference-order builds of the double-precision loops (exact.cuh): bit-identical to the reference's CPU solvers
int solve_real_x(Engine& E, const Operator<double>& A, int solver_id, double* m, const double* B, const double* lo, const double* hi,
	const lcgb200_para& para, size_t n, size_t next);
int solve_complex_x(Engine& E, const Operator<double2>& A, int solver_id, double2* m, const double2* B,
	const lcgb200_cpara& para, size_t n, size_t next);

const char* last_error();

}  // namespace lcgb200
