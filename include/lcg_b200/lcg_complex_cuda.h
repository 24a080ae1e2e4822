// lcg_b200/lcg_complex_cuda.h — C++ drop-in for the device helpers of liblcg's src/lib/lcg_complex_cuda.h (single and double
// precision complex).  All arrays are DEVICE arrays; asynchronous on the default stream; block size accepted and ignored.
//
//   clcg_smCcsr_get_diagonal / clcg_smZcsr_get_diagonal       lcg_complex_cuda.h:188,202  -> lcgb200_diagonal_of_csr
//   clcg_vecMvecC/Z_element_wise, clcg_vecDvecC/Z_element_wise lcg_complex_cuda.h:215-254 -> lcgb200_vec_elementwise (op 0 / 1)
//   clcg_vecC_conjugate / clcg_vecZ_conjugate                  lcg_complex_cuda.h:264,274  -> lcgb200_vec_elementwise (op 2)
//
// The host-side value helpers of the same header (clcg_Zsum, clcg_smZcoo_row2col, ...) are defined out of line in liblcg_dropin.so.
#ifndef LCG_B200_LCG_COMPLEX_CUDA_H
#define LCG_B200_LCG_COMPLEX_CUDA_H

#include <cuComplex.h>
#include "util.h"

inline void clcg_smCcsr_get_diagonal(const int* A_ptr, const int* A_col, const cuComplex* A_val, const int A_len, cuComplex* A_diag, int bk_size = 1024)
{
	(void)bk_size;
	lcgb200_diagonal_of_csr(LCGB200_COMPLEX_FLOAT, A_ptr, A_col, A_val, A_len, A_diag, nullptr);
}
inline void clcg_smZcsr_get_diagonal(const int* A_ptr, const int* A_col, const cuDoubleComplex* A_val, const int A_len, cuDoubleComplex* A_diag, int bk_size = 1024)
{
	(void)bk_size;
	lcgb200_diagonal_of_csr(LCGB200_COMPLEX, A_ptr, A_col, A_val, A_len, A_diag, nullptr);
}
inline void clcg_vecMvecC_element_wise(const cuComplex* a, const cuComplex* b, cuComplex* c, int n, int bk_size = 1024) { (void)bk_size; lcgb200_vec_elementwise(0, LCGB200_COMPLEX_FLOAT, a, b, c, n, nullptr); }
inline void clcg_vecMvecZ_element_wise(const cuDoubleComplex* a, const cuDoubleComplex* b, cuDoubleComplex* c, int n, int bk_size = 1024) { (void)bk_size; lcgb200_vec_elementwise(0, LCGB200_COMPLEX, a, b, c, n, nullptr); }
inline void clcg_vecDvecC_element_wise(const cuComplex* a, const cuComplex* b, cuComplex* c, int n, int bk_size = 1024) { (void)bk_size; lcgb200_vec_elementwise(1, LCGB200_COMPLEX_FLOAT, a, b, c, n, nullptr); }
inline void clcg_vecDvecZ_element_wise(const cuDoubleComplex* a, const cuDoubleComplex* b, cuDoubleComplex* c, int n, int bk_size = 1024) { (void)bk_size; lcgb200_vec_elementwise(1, LCGB200_COMPLEX, a, b, c, n, nullptr); }
inline void clcg_vecC_conjugate(const cuComplex* a, cuComplex* ca, int n, int bk_size = 1024) { (void)bk_size; lcgb200_vec_elementwise(2, LCGB200_COMPLEX_FLOAT, a, nullptr, ca, n, nullptr); }
inline void clcg_vecZ_conjugate(const cuDoubleComplex* a, cuDoubleComplex* ca, int n, int bk_size = 1024) { (void)bk_size; lcgb200_vec_elementwise(2, LCGB200_COMPLEX, a, nullptr, ca, n, nullptr); }

#endif  // LCG_B200_LCG_COMPLEX_CUDA_H
