// lcg_b200/lcg_cuda.h — C++ drop-in for liblcg's src/lib/lcg_cuda.h: the same callback typedefs and the same
// three entry points (names, argument order, default solver ids), forwarding to liblcgb200.so.
//
//   lcg_solver_cuda                 lcg_cuda.h:81-83     -> lcgb200_solver_cuda
//   lcg_solver_preconditioned_cuda  lcg_cuda.h:104-106   -> lcgb200_solver_preconditioned_cuda
//   lcg_solver_constrained_cuda     lcg_cuda.h:129-131   -> lcgb200_solver_constrained_cuda
//
// m and B are HOST arrays, exactly as the reference's implementation treats them (lcg_cuda.cu:110-111,210).
// Existing callbacks (cusparseSpMV inside Afp, sample8.cu:96-103) keep working unchanged.  To switch to the fused
// built-in operator pass  lcgb200_csr_ax / lcgb200_jacobi_mx  as Afp / Mfp and an lcgb200_csr_t as `instance`.
#ifndef LCG_B200_LCG_CUDA_H
#define LCG_B200_LCG_CUDA_H

#include <cublas_v2.h>
#include <cusparse_v2.h>
#include "util.h"

// lcg_cuda.h:45-46
typedef void (*lcg_axfunc_cuda_ptr)(void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax, const int n_size, const int nz_size);
// lcg_cuda.h:61-62 — m points at the solver's DEVICE copy of the current solution
typedef int (*lcg_progress_cuda_ptr)(void* instance, const lcg_float* m, const lcg_float converge,
	const lcg_para* param, const int n_size, const int nz_size, const int k);

inline int lcg_solver_cuda(lcg_axfunc_cuda_ptr Afp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B,
	const int n_size, const int nz_size, const lcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, lcg_solver_enum solver_id = LCG_CG)
{
	return lcgb200_solver_cuda(reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_progress_cuda_ptr>(Pfp),
		m, B, n_size, nz_size, param, instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

inline int lcg_solver_preconditioned_cuda(lcg_axfunc_cuda_ptr Afp, lcg_axfunc_cuda_ptr Mfp, lcg_progress_cuda_ptr Pfp,
	lcg_float* m, const lcg_float* B, const int n_size, const int nz_size, const lcg_para* param, void* instance,
	cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_solver_enum solver_id = LCG_PCG)
{
	return lcgb200_solver_preconditioned_cuda(reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Mfp),
		reinterpret_cast<lcgb200_progress_cuda_ptr>(Pfp), m, B, n_size, nz_size, param, instance,
		reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

inline int lcg_solver_constrained_cuda(lcg_axfunc_cuda_ptr Afp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B,
	const lcg_float* low, const lcg_float* hig, const int n_size, const int nz_size, const lcg_para* param, void* instance,
	cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_solver_enum solver_id = LCG_PG)
{
	return lcgb200_solver_constrained_cuda(reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_progress_cuda_ptr>(Pfp),
		m, B, low, hig, n_size, nz_size, param, instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

#endif  // LCG_B200_LCG_CUDA_H
