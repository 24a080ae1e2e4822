// lcg_b200/clcg_cudaf.h — C++ drop-in for liblcg's src/lib/clcg_cudaf.h (complex single precision, cuComplex):
//
//   clcg_solver_cuda (cuComplex overload)                 clcg_cudaf.h:81-83    -> lcgb200_csolver_cudaf
//   clcg_solver_preconditioned_cuda (cuComplex overload)  clcg_cudaf.h:103-105  -> lcgb200_csolver_preconditioned_cudaf
//
// m and B are HOST cuComplex arrays.  BICG and BICG_SYM / PCG, as in clcg_cudaf.cu.  Vectors and matrix values are stored as
// floats; dot products, norms and the iteration scalars are carried in double.  Built-in operator: lcgb200_csr_cax /
// lcgb200_jacobi_cmx with an lcgb200_csr_t created with LCGB200_COMPLEX_FLOAT (LCGB200_CSR_TRANSPOSE for CLCG_BICG).
#ifndef LCG_B200_CLCG_CUDAF_H
#define LCG_B200_CLCG_CUDAF_H

#include <cuComplex.h>
#include <cublas_v2.h>
#include <cusparse_v2.h>
#include "util.h"

// clcg_cudaf.h:45-46
typedef void (*clcg_axfunc_cudaf_ptr)(void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax, const int n_size, const int nz_size, cusparseOperation_t oper_t);
// clcg_cudaf.h:61-62
typedef int (*clcg_progress_cudaf_ptr)(void* instance, const cuComplex* m, const float converge,
	const clcg_para* param, const int n_size, const int nz_size, const int k);

inline int clcg_solver_cuda(clcg_axfunc_cudaf_ptr Afp, clcg_progress_cudaf_ptr Pfp, cuComplex* m, const cuComplex* B,
	const int n_size, const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, clcg_solver_enum solver_id = CLCG_BICG)
{
	return lcgb200_csolver_cudaf(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_cprogress_cudaf_ptr>(Pfp),
		m, B, n_size, nz_size, param, instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

inline int clcg_solver_preconditioned_cuda(clcg_axfunc_cudaf_ptr Afp, clcg_axfunc_cudaf_ptr Mfp, clcg_progress_cudaf_ptr Pfp,
	cuComplex* m, const cuComplex* B, const int n_size, const int nz_size, const clcg_para* param, void* instance,
	cublasHandle_t cub_handle, cusparseHandle_t cus_handle, clcg_solver_enum solver_id = CLCG_PCG)
{
	return lcgb200_csolver_preconditioned_cudaf(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Mfp),
		reinterpret_cast<lcgb200_cprogress_cudaf_ptr>(Pfp), m, B, n_size, nz_size, param, instance,
		reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

#endif  // LCG_B200_CLCG_CUDAF_H
