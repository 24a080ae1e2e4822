// lcg_b200/clcg_cuda.h — C++ drop-in for liblcg's src/lib/clcg_cuda.h (complex double):
//
//   clcg_solver_cuda                 clcg_cuda.h:81-83    -> lcgb200_csolver_cuda
//   clcg_solver_preconditioned_cuda  clcg_cuda.h:103-105  -> lcgb200_csolver_preconditioned_cuda
//
// m and B are HOST cuDoubleComplex arrays (clcg_cuda.cu:112-113,241).  Built-in operator: pass lcgb200_csr_cax /
// lcgb200_jacobi_cmx and an lcgb200_csr_t (created with LCGB200_COMPLEX; LCGB200_CSR_TRANSPOSE for CLCG_BICG).
// Besides BICG / BICG_SYM / PCG (all the reference's CUDA build offers) CGS, BICGSTAB and TFQMR are accepted.
// Residual definition: the reference CPU solver's (clcg.cpp:112-147) by default; lcgb200_set_complex_residual_mode(1)
// selects the reference CUDA solver's (clcg_cuda.cu:145-176).
#ifndef LCG_B200_CLCG_CUDA_H
#define LCG_B200_CLCG_CUDA_H

#include <cuComplex.h>
#include <cublas_v2.h>
#include <cusparse_v2.h>
#include "util.h"

// clcg_cuda.h:45-46
typedef void (*clcg_axfunc_cuda_ptr)(void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax, const int n_size, const int nz_size, cusparseOperation_t oper_t);
// clcg_cuda.h:61-62
typedef int (*clcg_progress_cuda_ptr)(void* instance, const cuDoubleComplex* m, const lcg_float converge,
	const clcg_para* param, const int n_size, const int nz_size, const int k);

inline int clcg_solver_cuda(clcg_axfunc_cuda_ptr Afp, clcg_progress_cuda_ptr Pfp, cuDoubleComplex* m, const cuDoubleComplex* B,
	const int n_size, const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, clcg_solver_enum solver_id = CLCG_BICG)
{
	return lcgb200_csolver_cuda(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_cprogress_cuda_ptr>(Pfp),
		m, B, n_size, nz_size, param, instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

inline int clcg_solver_preconditioned_cuda(clcg_axfunc_cuda_ptr Afp, clcg_axfunc_cuda_ptr Mfp, clcg_progress_cuda_ptr Pfp,
	cuDoubleComplex* m, const cuDoubleComplex* B, const int n_size, const int nz_size, const clcg_para* param, void* instance,
	cublasHandle_t cub_handle, cusparseHandle_t cus_handle, clcg_solver_enum solver_id = CLCG_PCG)
{
	return lcgb200_csolver_preconditioned_cuda(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Mfp),
		reinterpret_cast<lcgb200_cprogress_cuda_ptr>(Pfp), m, B, n_size, nz_size, param, instance,
		reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

#endif  // LCG_B200_CLCG_CUDA_H
