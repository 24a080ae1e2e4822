// lcg_b200/clcg.h — C++ drop-in for the HOST-callback complex API of liblcg's src/lib/clcg.h:
//   clcg_solver  clcg.h:74-76  -> lcgb200_csolver   (default CLCG_BICG, as there)
// lcg_complex is std::complex<double> (the reference's LibLCG_STD_COMPLEX build, lcg_complex.h:33); the Ax callback
// receives (layout, conjugate) exactly like clcg.h:40-41 — BiCG asks for (MatTranspose, Conjugate), clcg.cpp:188.
// Built-in operator: lcgb200_csr_cax_host + an lcgb200_csr_t created with LCGB200_COMPLEX (| LCGB200_CSR_TRANSPOSE for BICG).
#ifndef LCG_B200_CLCG_H
#define LCG_B200_CLCG_H

#include <complex>
#include "util.h"

typedef std::complex<lcg_float> lcg_complex;
enum lcg_matrix_e { MatNormal, MatTranspose };      // algebra.h:31-35
enum clcg_complex_e { NonConjugate, Conjugate };    // algebra.h:40-44

typedef void (*clcg_axfunc_ptr)(void* instance, const lcg_complex* x, lcg_complex* prod_Ax, const int x_size, lcg_matrix_e layout, clcg_complex_e conjugate);
typedef int (*clcg_progress_ptr)(void* instance, const lcg_complex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int k);

inline int clcg_solver(clcg_axfunc_ptr Afp, clcg_progress_ptr Pfp, lcg_complex* m, const lcg_complex* B, const int n_size,
	const clcg_para* param, void* instance, clcg_solver_enum solver_id = CLCG_BICG)
{
	// enum arguments are passed as int in both ABIs; std::complex<double> is two interleaved doubles
	return lcgb200_csolver(reinterpret_cast<lcgb200_caxfunc_ptr>(Afp), reinterpret_cast<lcgb200_cprogress_ptr>(Pfp), m, B, n_size, param, instance,
		static_cast<int>(solver_id));
}

#endif  // LCG_B200_CLCG_H
