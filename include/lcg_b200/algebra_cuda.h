// lcg_b200/algebra_cuda.h — C++ drop-in for liblcg's src/lib/algebra_cuda.h: the device helpers the reference's samples build
// their Jacobi preconditioner callbacks from (sample10.cu:117,193).  All arrays are DEVICE arrays; the calls are asynchronous on
// the default stream like the reference's kernel launches; the block-size argument is accepted and ignored.
//
//   lcg_set2box_cuda            algebra_cuda.h:45-46  -> lcgb200_set2box
//   lcg_smDcsr_get_diagonal     algebra_cuda.h:58     -> lcgb200_diagonal_of_csr
//   lcg_vecMvecD_element_wise   algebra_cuda.h:71     -> lcgb200_vec_elementwise (op 0)
//   lcg_vecDvecD_element_wise   algebra_cuda.h:84     -> lcgb200_vec_elementwise (op 1)
#ifndef LCG_B200_ALGEBRA_CUDA_H
#define LCG_B200_ALGEBRA_CUDA_H

#include "util.h"

inline void lcg_set2box_cuda(const lcg_float* low, const lcg_float* hig, lcg_float* a, int n, bool low_bound = true, bool hig_bound = true)
{
	(void)low_bound; (void)hig_bound;   // closed or open bounds clamp to the same values (algebra_cuda.cu:26-38)
	lcgb200_set2box(low, hig, a, n, nullptr);
}
inline void lcg_smDcsr_get_diagonal(const int* A_ptr, const int* A_col, const lcg_float* A_val, const int A_len, lcg_float* A_diag, int bk_size = 1024)
{
	(void)bk_size;
	lcgb200_diagonal_of_csr(LCGB200_REAL, A_ptr, A_col, A_val, A_len, A_diag, nullptr);
}
inline void lcg_vecMvecD_element_wise(const lcg_float* a, const lcg_float* b, lcg_float* c, int n, int bk_size = 1024)
{
	(void)bk_size;
	lcgb200_vec_elementwise(0, LCGB200_REAL, a, b, c, n, nullptr);
}
inline void lcg_vecDvecD_element_wise(const lcg_float* a, const lcg_float* b, lcg_float* c, int n, int bk_size = 1024)
{
	(void)bk_size;
	lcgb200_vec_elementwise(1, LCGB200_REAL, a, b, c, n, nullptr);
}

#endif  // LCG_B200_ALGEBRA_CUDA_H
