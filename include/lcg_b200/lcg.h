// lcg_b200/lcg.h — C++ drop-in for the HOST-callback API of liblcg's src/lib/lcg.h:
//   lcg_solver                 lcg.h:71-72    -> lcgb200_solver                 (default LCG_CGS, as there)
//   lcg_solver_preconditioned  lcg.h:90-91    -> lcgb200_solver_preconditioned
//   lcg_solver_constrained     lcg.h:111-113  -> lcgb200_solver_constrained
// Callback typedefs lcg_axfunc_ptr / lcg_progress_ptr as in lcg.h:37-38,53-54.  Existing host callbacks keep working
// (they run on the host; the vectors are staged over PCIe per call).  Pass lcgb200_csr_ax_host / lcgb200_jacobi_mx_host
// and an lcgb200_csr_t as `instance` to run the whole solve on the GPU's fused built-in operator.
#ifndef LCG_B200_LCG_H
#define LCG_B200_LCG_H

#include "util.h"

typedef void (*lcg_axfunc_ptr)(void* instance, const lcg_float* x, lcg_float* prod_Ax, const int n_size);
typedef int (*lcg_progress_ptr)(void* instance, const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int k);

inline int lcg_solver(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size,
	const lcg_para* param, void* instance, lcg_solver_enum solver_id = LCG_CGS)
{
	return lcgb200_solver(Afp, Pfp, m, B, n_size, param, instance, static_cast<int>(solver_id));
}

inline int lcg_solver_preconditioned(lcg_axfunc_ptr Afp, lcg_axfunc_ptr Mfp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B,
	const int n_size, const lcg_para* param, void* instance, lcg_solver_enum solver_id = LCG_PCG)
{
	return lcgb200_solver_preconditioned(Afp, Mfp, Pfp, m, B, n_size, param, instance, static_cast<int>(solver_id));
}

inline int lcg_solver_constrained(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const lcg_float* low,
	const lcg_float* hig, const int n_size, const lcg_para* param, void* instance, lcg_solver_enum solver_id = LCG_PG)
{
	return lcgb200_solver_constrained(Afp, Pfp, m, B, low, hig, n_size, param, instance, static_cast<int>(solver_id));
}

#endif  // LCG_B200_LCG_H
