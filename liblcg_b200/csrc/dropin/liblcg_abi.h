// liblcg_abi.h — the C++ ABI of liblcg (YiZhangCUG/liblcg, src/lib) restated for liblcg_dropin.so: type names, enum
// names, struct layouts, function signatures and class layouts exactly as a program compiled against the reference's OWN
// headers expects them, so that such a program links against (or is re-pointed at) our library without recompiling.
// Mangled names depend on these spellings: `struct lcg_para` (not a typedef of another struct), `enum lcg_solver_enum`,
// `std::complex<double>` for lcg_complex (config.h:3 LibLCG_STD_COMPLEX), cuBLAS/cuSPARSE handle types from the CUDA headers.
//   util.h:32-153,187-306   algebra.h:31-217   lcg_complex.h:190-340   lcg.h:37-169   clcg.h:40-76
//   lcg_cuda.h:45-131   clcg_cuda.h:45-105   solver.h:32-283   solver_cuda.h:35-207,380-541
// Everything here is declared, nothing is inline: the definitions live in liblcg_dropin.cpp and forward to the C ABI.
#pragma once
#include <complex>
#include <string>
#include <cublas_v2.h>
#include <cusparse_v2.h>
#include <cuComplex.h>

typedef double lcg_float;
typedef std::complex<lcg_float> lcg_complex;
enum lcg_matrix_e { MatNormal, MatTranspose };
enum clcg_complex_e { NonConjugate, Conjugate };

enum lcg_solver_enum { LCG_CG, LCG_PCG, LCG_CGS, LCG_BICGSTAB, LCG_BICGSTAB2, LCG_PG, LCG_SPG };
enum lcg_return_enum {
	LCG_SUCCESS = 0, LCG_CONVERGENCE = 0, LCG_STOP, LCG_ALREADY_OPTIMIZIED,
	LCG_UNKNOWN_ERROR = -1024, LCG_INVILAD_VARIABLE_SIZE, LCG_INVILAD_MAX_ITERATIONS, LCG_INVILAD_EPSILON, LCG_INVILAD_RESTART_EPSILON,
	LCG_REACHED_MAX_ITERATIONS, LCG_NULL_PRECONDITION_MATRIX, LCG_NAN_VALUE, LCG_INVALID_POINTER, LCG_INVALID_LAMBDA, LCG_INVALID_SIGMA,
	LCG_INVALID_BETA, LCG_INVALID_MAXIM, LCG_SIZE_NOT_MATCH
};
struct lcg_para { int max_iterations; lcg_float epsilon; int abs_diff; lcg_float restart_epsilon; lcg_float step; lcg_float sigma; lcg_float beta; int maxi_m; };
enum clcg_solver_enum { CLCG_BICG, CLCG_BICG_SYM, CLCG_CGS, CLCG_BICGSTAB, CLCG_TFQMR, CLCG_PCG, CLCG_PBICG };
enum clcg_return_enum {
	CLCG_SUCCESS = 0, CLCG_CONVERGENCE = 0, CLCG_STOP, CLCG_ALREADY_OPTIMIZIED,
	CLCG_UNKNOWN_ERROR = -1024, CLCG_INVILAD_VARIABLE_SIZE, CLCG_INVILAD_MAX_ITERATIONS, CLCG_INVILAD_EPSILON, CLCG_REACHED_MAX_ITERATIONS,
	CLCG_NAN_VALUE, CLCG_INVALID_POINTER, CLCG_SIZE_NOT_MATCH, CLCG_UNKNOWN_SOLVER
};
struct clcg_para { int max_iterations; lcg_float epsilon; int abs_diff; };

// ---- util.h
lcg_para lcg_default_parameters();
lcg_solver_enum lcg_select_solver(std::string slr_char);
void lcg_error_str(int er_index, bool er_throw);
clcg_para clcg_default_parameters();
clcg_solver_enum clcg_select_solver(std::string slr_char);
void clcg_error_str(int er_index, bool er_throw);

// ---- algebra.h / lcg_complex.h: the small host helpers every sample program uses around a solve
lcg_float lcg_abs(lcg_float a);
lcg_float lcg_max(lcg_float a, lcg_float b);
lcg_float lcg_min(lcg_float a, lcg_float b);
lcg_float lcg_set2box(lcg_float low, lcg_float hig, lcg_float a, bool low_bound, bool hig_bound);
lcg_float* lcg_malloc(int n);
lcg_float** lcg_malloc(int m, int n);
void lcg_free(lcg_float* x);
void lcg_free(lcg_float** x, int m);
void lcg_vecset(lcg_float* a, lcg_float b, int size);
void lcg_vecset(lcg_float** a, lcg_float b, int m, int n);
void lcg_vecrnd(lcg_float* a, lcg_float l, lcg_float h, int size);
void lcg_vecrnd(lcg_float** a, lcg_float l, lcg_float h, int m, int n);
double lcg_squaredl2norm(lcg_float* a, int n);
void lcg_dot(lcg_float& ret, const lcg_float* a, const lcg_float* b, int size);
void lcg_matvec(lcg_float** A, const lcg_float* x, lcg_float* Ax, int m_size, int n_size, lcg_matrix_e layout);
void lcg_matvec_coo(const int* row, const int* col, const lcg_float* Mat, const lcg_float* V, lcg_float* p, int M, int N, int nz_size, bool pre_position);
lcg_complex* clcg_malloc(int n);
lcg_complex** clcg_malloc(int m, int n);
void clcg_free(lcg_complex* x);
void clcg_free(lcg_complex** x, int m);
void clcg_vecset(lcg_complex* a, lcg_complex b, int size);
void clcg_vecset(lcg_complex** a, lcg_complex b, int m, int n);
void clcg_set(lcg_complex* a, lcg_float r, lcg_float i);
lcg_float clcg_square(const lcg_complex* a);
lcg_float clcg_module(const lcg_complex* a);
lcg_complex clcg_conjugate(const lcg_complex* a);
void clcg_vecrnd(lcg_complex* a, lcg_complex l, lcg_complex h, int size);
void clcg_vecrnd(lcg_complex** a, lcg_complex l, lcg_complex h, int m, int n);
void clcg_dot(lcg_complex& ret, const lcg_complex* a, const lcg_complex* b, int size);
void clcg_inner(lcg_complex& ret, const lcg_complex* a, const lcg_complex* b, int size);
void clcg_matvec(lcg_complex** A, const lcg_complex* x, lcg_complex* Ax, int m_size, int n_size, lcg_matrix_e layout, clcg_complex_e conjugate);

// ---- algebra_cuda.h:45-84 / lcg_complex_cuda.h:39-274: the device helpers the samples build their Jacobi callbacks from
// (sample10.cu:117,193) and the small host helpers around cuComplex values
void lcg_set2box_cuda(const lcg_float* low, const lcg_float* hig, lcg_float* a, int n, bool low_bound, bool hig_bound);
void lcg_smDcsr_get_diagonal(const int* A_ptr, const int* A_col, const lcg_float* A_val, const int A_len, lcg_float* A_diag, int bk_size);
void lcg_vecMvecD_element_wise(const lcg_float* a, const lcg_float* b, lcg_float* c, int n, int bk_size);
void lcg_vecDvecD_element_wise(const lcg_float* a, const lcg_float* b, lcg_float* c, int n, int bk_size);
lcg_complex cuda2lcg_complex(cuDoubleComplex a);
cuDoubleComplex lcg2cuda_complex(lcg_complex a);
cuDoubleComplex* clcg_malloc_cuda(size_t n);
void clcg_free_cuda(cuDoubleComplex* x);
void clcg_vecset_cuda(cuDoubleComplex* a, cuDoubleComplex b, size_t size);
cuComplex clcg_Cscale(float s, cuComplex a);
cuComplex clcg_Csum(cuComplex a, cuComplex b);
cuComplex clcg_Cdiff(cuComplex a, cuComplex b);
cuComplex clcg_Csqrt(cuComplex a);
cuDoubleComplex clcg_Zscale(lcg_float s, cuDoubleComplex a);
cuDoubleComplex clcg_Zsum(cuDoubleComplex a, cuDoubleComplex b);
cuDoubleComplex clcg_Zdiff(cuDoubleComplex a, cuDoubleComplex b);
cuDoubleComplex clcg_Zsqrt(cuDoubleComplex a);
void clcg_smCcoo_row2col(const int* A_row, const int* A_col, const cuComplex* A, int N, int nz, int* Ac_row, int* Ac_col, cuComplex* Ac_val);
void clcg_smZcoo_row2col(const int* A_row, const int* A_col, const cuDoubleComplex* A, int N, int nz, int* Ac_row, int* Ac_col, cuDoubleComplex* Ac_val);
void clcg_smCcsr_get_diagonal(const int* A_ptr, const int* A_col, const cuComplex* A_val, const int A_len, cuComplex* A_diag, int bk_size);
void clcg_smZcsr_get_diagonal(const int* A_ptr, const int* A_col, const cuDoubleComplex* A_val, const int A_len, cuDoubleComplex* A_diag, int bk_size);
void clcg_vecMvecC_element_wise(const cuComplex* a, const cuComplex* b, cuComplex* c, int n, int bk_size);
void clcg_vecMvecZ_element_wise(const cuDoubleComplex* a, const cuDoubleComplex* b, cuDoubleComplex* c, int n, int bk_size);
void clcg_vecDvecC_element_wise(const cuComplex* a, const cuComplex* b, cuComplex* c, int n, int bk_size);
void clcg_vecDvecZ_element_wise(const cuDoubleComplex* a, const cuDoubleComplex* b, cuDoubleComplex* c, int n, int bk_size);
void clcg_vecC_conjugate(const cuComplex* a, cuComplex* ca, int n, int bk_size);
void clcg_vecZ_conjugate(const cuDoubleComplex* a, cuDoubleComplex* ca, int n, int bk_size);

// ---- preconditioner.h / preconditioner_cuda.h: IC(0) of a row-sorted COO matrix and the COO triangular solves (host code)
void lcg_incomplete_Cholesky_half_buffsize_coo(const int* row, const int* col, int nz_size, int* lnz_size);
void lcg_incomplete_Cholesky_half_coo(const int* row, const int* col, const lcg_float* val, int N, int nz_size, int lnz_size, int* IC_row, int* IC_col,
	lcg_float* IC_val);
void lcg_incomplete_Cholesky_full_coo(const int* row, const int* col, const lcg_float* val, int N, int nz_size, int* IC_row, int* IC_col, lcg_float* IC_val);
void lcg_solve_upper_triangle_coo(const int* row, const int* col, const lcg_float* U, const lcg_float* B, lcg_float* x, int N, int nz_size);
void lcg_solve_lower_triangle_coo(const int* row, const int* col, const lcg_float* L, const lcg_float* B, lcg_float* x, int N, int nz_size);
bool lcg_full_rank_coo(const int* row, const int* col, const lcg_float* M, int N, int nz_size);
void clcg_incomplete_Cholesky_cuda_half_buffsize(const int* row, const int* col, int nz_size, int* lnz_size);
void clcg_incomplete_Cholesky_cuda_half(const int* row, const int* col, const cuComplex* val, int N, int nz_size, int lnz_size, int* IC_row, int* IC_col,
	cuComplex* IC_val);
void clcg_incomplete_Cholesky_cuda_half(const int* row, const int* col, const cuDoubleComplex* val, int N, int nz_size, int lnz_size, int* IC_row,
	int* IC_col, cuDoubleComplex* IC_val);
void clcg_incomplete_Cholesky_cuda_full(const int* row, const int* col, const cuDoubleComplex* val, int N, int nz_size, int* IC_row, int* IC_col,
	cuDoubleComplex* IC_val);

// ---- lcg.h / clcg.h: host-callback solvers
typedef void (*lcg_axfunc_ptr)(void* instance, const lcg_float* x, lcg_float* prod_Ax, const int n_size);
typedef int (*lcg_progress_ptr)(void* instance, const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int k);
int lcg_solver(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const lcg_para* param, void* instance,
	lcg_solver_enum solver_id);
int lcg_solver_preconditioned(lcg_axfunc_ptr Afp, lcg_axfunc_ptr Mfp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size,
	const lcg_para* param, void* instance, lcg_solver_enum solver_id);
int lcg_solver_constrained(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const lcg_float* low, const lcg_float* hig,
	const int n_size, const lcg_para* param, void* instance, lcg_solver_enum solver_id);
int lcg(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const lcg_para* param, void* instance,
	lcg_float* Gk, lcg_float* Dk, lcg_float* ADk);
int lcgs(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const lcg_para* param, void* instance,
	lcg_float* RK, lcg_float* R0T, lcg_float* PK, lcg_float* AX, lcg_float* UK, lcg_float* QK, lcg_float* WK);
typedef void (*clcg_axfunc_ptr)(void* instance, const lcg_complex* x, lcg_complex* prod_Ax, const int x_size, lcg_matrix_e layout, clcg_complex_e conjugate);
typedef int (*clcg_progress_ptr)(void* instance, const lcg_complex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int k);
int clcg_solver(clcg_axfunc_ptr Afp, clcg_progress_ptr Pfp, lcg_complex* m, const lcg_complex* B, const int n_size, const clcg_para* param,
	void* instance, clcg_solver_enum solver_id);

// ---- lcg_cuda.h / clcg_cuda.h: the CUDA entry points
typedef void (*lcg_axfunc_cuda_ptr)(void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x,
	cusparseDnVecDescr_t prod_Ax, const int n_size, const int nz_size);
typedef int (*lcg_progress_cuda_ptr)(void* instance, const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size,
	const int nz_size, const int k);
int lcg_solver_cuda(lcg_axfunc_cuda_ptr Afp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const int nz_size,
	const lcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_solver_enum solver_id);
int lcg_solver_preconditioned_cuda(lcg_axfunc_cuda_ptr Afp, lcg_axfunc_cuda_ptr Mfp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B,
	const int n_size, const int nz_size, const lcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	lcg_solver_enum solver_id);
int lcg_solver_constrained_cuda(lcg_axfunc_cuda_ptr Afp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B, const lcg_float* low,
	const lcg_float* hig, const int n_size, const int nz_size, const lcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, lcg_solver_enum solver_id);
typedef void (*clcg_axfunc_cuda_ptr)(void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x,
	cusparseDnVecDescr_t prod_Ax, const int n_size, const int nz_size, cusparseOperation_t oper_t);
typedef int (*clcg_progress_cuda_ptr)(void* instance, const cuDoubleComplex* m, const lcg_float converge, const clcg_para* param, const int n_size,
	const int nz_size, const int k);
int clcg_solver_cuda(clcg_axfunc_cuda_ptr Afp, clcg_progress_cuda_ptr Pfp, cuDoubleComplex* m, const cuDoubleComplex* B, const int n_size,
	const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, clcg_solver_enum solver_id);
int clcg_solver_preconditioned_cuda(clcg_axfunc_cuda_ptr Afp, clcg_axfunc_cuda_ptr Mfp, clcg_progress_cuda_ptr Pfp, cuDoubleComplex* m,
	const cuDoubleComplex* B, const int n_size, const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, clcg_solver_enum solver_id);

// ---- clcg_cudaf.h: the cuComplex overloads
typedef void (*clcg_axfunc_cudaf_ptr)(void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x,
	cusparseDnVecDescr_t prod_Ax, const int n_size, const int nz_size, cusparseOperation_t oper_t);
typedef int (*clcg_progress_cudaf_ptr)(void* instance, const cuComplex* m, const float converge, const clcg_para* param, const int n_size,
	const int nz_size, const int k);
int clcg_solver_cuda(clcg_axfunc_cudaf_ptr Afp, clcg_progress_cudaf_ptr Pfp, cuComplex* m, const cuComplex* B, const int n_size, const int nz_size,
	const clcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, clcg_solver_enum solver_id);
int clcg_solver_preconditioned_cuda(clcg_axfunc_cudaf_ptr Afp, clcg_axfunc_cudaf_ptr Mfp, clcg_progress_cudaf_ptr Pfp, cuComplex* m, const cuComplex* B,
	const int n_size, const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	clcg_solver_enum solver_id);

// ---- solver.h / solver_cuda.h: the class wrappers.  Data members and virtual functions in the reference's order (the
// object layout and the vtable are part of the ABI); the static trampolines are inline in the reference's header and are
// therefore compiled into the caller, not into the library.
class LCG_Solver {
protected:
	lcg_para param_; unsigned int inter_; bool silent_;
public:
	LCG_Solver();
	virtual ~LCG_Solver() {}
	virtual void AxProduct(const lcg_float* a, lcg_float* b, const int num) = 0;
	virtual void MxProduct(const lcg_float* a, lcg_float* b, const int num) = 0;
	virtual int Progress(const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int k);
	void silent();
	void set_report_interval(unsigned int inter);
	void set_lcg_parameter(const lcg_para& in_param);
	void Minimize(lcg_float* m, const lcg_float* b, int x_size, lcg_solver_enum solver_id, bool verbose, bool er_throw);
	void MinimizePreconditioned(lcg_float* m, const lcg_float* b, int x_size, lcg_solver_enum solver_id, bool verbose, bool er_throw);
	void MinimizeConstrained(lcg_float* m, const lcg_float* b, const lcg_float* low, const lcg_float* hig, int x_size, lcg_solver_enum solver_id,
		bool verbose, bool er_throw);
};
class CLCG_Solver {
protected:
	clcg_para param_; unsigned int inter_; bool silent_;
public:
	CLCG_Solver();
	virtual ~CLCG_Solver() {}
	virtual void AxProduct(const lcg_complex* x, lcg_complex* prod_Ax, const int x_size, lcg_matrix_e layout, clcg_complex_e conjugate) = 0;
	virtual int Progress(const lcg_complex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int k);
	void silent();
	void set_report_interval(unsigned int inter);
	void set_clcg_parameter(const clcg_para& in_param);
	void Minimize(lcg_complex* m, const lcg_complex* b, int x_size, clcg_solver_enum solver_id, bool verbose, bool er_throw);
};
class LCG_CUDA_Solver {
protected:
	lcg_para param_; unsigned int inter_; bool silent_;
public:
	LCG_CUDA_Solver();
	virtual ~LCG_CUDA_Solver() {}
	virtual void AxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax,
		const int n_size, const int nz_size) = 0;
	virtual void MxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Mx,
		const int n_size, const int nz_size) = 0;
	virtual int Progress(const lcg_float* m, const lcg_float converge, const lcg_para* param, const int n_size, const int nz_size, const int k);
	void silent();
	void set_report_interval(unsigned int inter);
	void set_lcg_parameter(const lcg_para& in_param);
	void Minimize(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_float* x, lcg_float* b, const int n_size, const int nz_size,
		lcg_solver_enum solver_id, bool verbose, bool er_throw);
	void MinimizePreconditioned(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_float* x, lcg_float* b, const int n_size,
		const int nz_size, lcg_solver_enum solver_id, bool verbose, bool er_throw);
	void MinimizeConstrained(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_float* x, const lcg_float* b, const lcg_float* low,
		const lcg_float* hig, const int n_size, const int nz_size, lcg_solver_enum solver_id, bool verbose, bool er_throw);
};
class CLCG_CUDA_Solver {
protected:
	clcg_para param_; unsigned int inter_; bool silent_;
public:
	CLCG_CUDA_Solver();
	virtual ~CLCG_CUDA_Solver() {}
	virtual void AxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual void MxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Mx,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual int Progress(const cuDoubleComplex* m, const lcg_float converge, const clcg_para* param, const int n_size, const int nz_size, const int k);
	void silent();
	void set_report_interval(unsigned int inter);
	void set_clcg_parameter(const clcg_para& in_param);
	void Minimize(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuDoubleComplex* x, cuDoubleComplex* b, const int n_size,
		const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw);
	void MinimizePreconditioned(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuDoubleComplex* x, cuDoubleComplex* b, const int n_size,
		const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw);
};
class CLCG_CUDAF_Solver {
protected:
	clcg_para param_; unsigned int inter_; bool silent_;
public:
	CLCG_CUDAF_Solver();
	virtual ~CLCG_CUDAF_Solver() {}
	virtual void AxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Ax,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual void MxProduct(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cusparseDnVecDescr_t x, cusparseDnVecDescr_t prod_Mx,
		const int n_size, const int nz_size, cusparseOperation_t oper_t) = 0;
	virtual int Progress(const cuComplex* m, const float converge, const clcg_para* param, const int n_size, const int nz_size, const int k);
	void silent();
	void set_report_interval(unsigned int inter);
	void set_clcg_parameter(const clcg_para& in_param);
	void Minimize(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuComplex* x, cuComplex* b, const int n_size, const int nz_size,
		clcg_solver_enum solver_id, bool verbose, bool er_throw);
	void MinimizePreconditioned(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuComplex* x, cuComplex* b, const int n_size,
		const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw);
};
