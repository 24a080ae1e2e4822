/*
 * lcgb200.h — C ABI of liblcgb200.so, the B200-native (sm_100a) replacement for the iteration loops of
 * liblcg's CUDA solvers.  Plain pointers and sizes only; no C++ types, no torch types.
 *
 * Every entry point cites the reference interface it replaces (paths under YiZhangCUG/liblcg src/lib).
 * The C++ drop-in headers in include/lcg_b200/ (lcg_cuda.h, clcg_cuda.h, lcg.h, clcg.h) re-declare the
 * reference's own function names (with their default arguments) as inline forwards to these symbols.
 *
 * Two ways in:
 *   1. Reference-shaped calls (lcgb200_solver_cuda & co.): same arguments, same meaning, same return codes
 *      as lcg_solver_cuda & co.  m and B are HOST arrays (the reference copies them itself,
 *      lcg_cuda.cu:110-111,210).  Any user Ax / Mx callback works (generic path: the user's SpMV, then our
 *      fused vector kernels).  Passing the exported sentinel callbacks lcgb200_csr_ax / lcgb200_jacobi_mx
 *      (lcgb200_csr_cax / lcgb200_jacobi_cmx for complex) with `instance` = an lcgb200_csr_t switches to the
 *      built-in fused CSR operator: SpMV fused with its dot products, no cuSPARSE/cuBLAS anywhere.
 *   2. Handle-shaped calls (lcgb200_solve / lcgb200_csolve): the same engine for callers whose vectors
 *      already live on the device, plus iteration statistics.
 */
#ifndef LCGB200_H
#define LCGB200_H

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque CUDA library types, declared exactly as cublas_v2.h / cusparse.h declare them so that this header
 * can be used with or without those headers. */
struct cublasContext;
struct cusparseContext;
struct cusparseDnVecDescr;
typedef struct cublasContext* lcgb200_cublas_t;       /* == cublasHandle_t */
typedef struct cusparseContext* lcgb200_cusparse_t;   /* == cusparseHandle_t */
typedef struct cusparseDnVecDescr* lcgb200_dnvec_t;   /* == cusparseDnVecDescr_t */

/* ---- parameter blocks: byte-identical to the reference (util.h:95-148, util.h:247-273) ---- */
typedef struct lcgb200_para {
	int max_iterations;      /* 0 = until convergence */
	double epsilon;          /* (0,1); test is on SQUARED norms: |r|^2/max(|m|^2,1) <= eps (lcg.cpp:208-209) */
	int abs_diff;            /* 1: sqrt(|r|^2)/n <= eps */
	double restart_epsilon;  /* BICGSTAB2 */
	double step;             /* PG / SPG initial step */
	double sigma;            /* SPG */
	double beta;             /* SPG */
	int maxi_m;              /* SPG history length */
} lcgb200_para;

typedef struct lcgb200_cpara {
	int max_iterations;
	double epsilon;
	int abs_diff;
} lcgb200_cpara;

/* ---- solver ids (util.h:32-64, util.h:187-221) ---- */
enum { LCGB200_CG = 0, LCGB200_PCG, LCGB200_CGS, LCGB200_BICGSTAB, LCGB200_BICGSTAB2, LCGB200_PG, LCGB200_SPG };
enum { LCGB200_CBICG = 0, LCGB200_CBICG_SYM, LCGB200_CCGS, LCGB200_CBICGSTAB, LCGB200_CTFQMR, LCGB200_CPCG, LCGB200_CPBICG };

/* ---- return codes: the exact integers of lcg_return_enum / clcg_return_enum (util.h:69-90, 226-242) ---- */
enum {
	LCGB200_CONVERGENCE = 0, LCGB200_STOP = 1, LCGB200_ALREADY_OPTIMIZIED = 2,
	LCGB200_UNKNOWN_ERROR = -1024, LCGB200_INVILAD_VARIABLE_SIZE = -1023, LCGB200_INVILAD_MAX_ITERATIONS = -1022,
	LCGB200_INVILAD_EPSILON = -1021, LCGB200_INVILAD_RESTART_EPSILON = -1020, LCGB200_REACHED_MAX_ITERATIONS = -1019,
	LCGB200_NULL_PRECONDITION_MATRIX = -1018, LCGB200_NAN_VALUE = -1017, LCGB200_INVALID_POINTER = -1016,
	LCGB200_INVALID_LAMBDA = -1015, LCGB200_INVALID_SIGMA = -1014, LCGB200_INVALID_BETA = -1013,
	LCGB200_INVALID_MAXIM = -1012, LCGB200_SIZE_NOT_MATCH = -1011,
	/* complex enum: same names shifted (util.h:226-242).  The complex solvers return the REAL enum's
	 * LCG_REACHED_MAX_ITERATIONS (-1019, clcg.cpp:164 / clcg_cuda.cu:193) and so do we. */
	LCGB200_C_REACHED_MAX_ITERATIONS = -1020, LCGB200_C_NAN_VALUE = -1019, LCGB200_C_INVALID_POINTER = -1018,
	LCGB200_C_SIZE_NOT_MATCH = -1017, LCGB200_C_UNKNOWN_SOLVER = -1016
};

/* ---- callback typedefs: source-compatible with lcg_cuda.h:45-46,61-62 and clcg_cuda.h:45-46,61-62 ---- */
typedef void (*lcgb200_axfunc_cuda_ptr)(void* instance, lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle,
	lcgb200_dnvec_t x, lcgb200_dnvec_t prod_Ax, const int n_size, const int nz_size);
typedef int (*lcgb200_progress_cuda_ptr)(void* instance, const double* m_dev, const double converge,
	const lcgb200_para* param, const int n_size, const int nz_size, const int k);
/* oper_t is a cusparseOperation_t: 0 = N, 1 = T, 2 = H */
typedef void (*lcgb200_caxfunc_cuda_ptr)(void* instance, lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle,
	lcgb200_dnvec_t x, lcgb200_dnvec_t prod_Ax, const int n_size, const int nz_size, int oper_t);
/* m_dev points at n_size cuDoubleComplex (interleaved re,im doubles) */
typedef int (*lcgb200_cprogress_cuda_ptr)(void* instance, const void* m_dev, const double converge,
	const lcgb200_cpara* param, const int n_size, const int nz_size, const int k);
/* host-side (CPU API) callbacks: lcg.h:37-38,53-54 */
typedef void (*lcgb200_axfunc_ptr)(void* instance, const double* x, double* prod_Ax, const int n_size);
typedef int (*lcgb200_progress_ptr)(void* instance, const double* m, const double converge,
	const lcgb200_para* param, const int n_size, const int k);

/* =====================================================================================================
 * Built-in CSR operator (new; the reference leaves the matrix to the caller — sample8.cu:80-103)
 * ===================================================================================================== */
typedef struct lcgb200_csr_s* lcgb200_csr_t;

enum { LCGB200_REAL = 0, LCGB200_COMPLEX = 1,   /* double / cuDoubleComplex (interleaved re, im doubles) */
       LCGB200_COMPLEX_FLOAT = 2 };              /* cuComplex (interleaved floats): the clcg_cudaf.h entry points */
enum { LCGB200_HOST = 0, LCGB200_DEVICE = 1 };
enum {
	LCGB200_CSR_TRANSPOSE = 1,   /* also store A^T (needed by complex BiCG's A^H d2, clcg.cpp:188) */
	LCGB200_CSR_COMPRESS = 4,    /* real operators: if the matrix has <= 256 distinct values and <= 256 distinct (col - row) offsets
	                                 (constant-coefficient stencils and their row blocks), keep a second copy as one 16-bit code per
	                                 entry + two dictionaries and stream THAT in the SpMV (2 bytes per non-zero instead of 12); if in
	                                 addition the rows fall into <= 253 distinct patterns, keep one pattern id per ROW and stream only
	                                 that.  Same entries; the row sums are accumulated in another order than on the plain copy (y
	                                 agrees to rounding).  Silently stays uncompressed when the matrix does not fit (lcgb200_csr_format). */
	LCGB200_CSR_IC0 = 8,         /* zero-fill incomplete Cholesky of the (symmetric, square, unpartitioned) matrix at creation: the
	                                 reference's sequential algorithm on the host (lcg_incomplete_Cholesky_half_coo, preconditioner.cpp:33-160;
	                                 clcg_incomplete_Cholesky_cuda_half, preconditioner_cuda.cu:40-270; complex: L L^T, unconjugated),
	                                 bit-identical factor; L, L^T and their level orders are kept on the device so that lcgb200_ic0_mx /
	                                 lcgb200_ic0_cmx apply z = L^-T L^-1 r by two level-ordered triangular solves (what the reference's
	                                 samples do with cusparseSpSV in their Mx callback, sample12.cu:95-105) */
	LCGB200_CSR_JACOBI = 2       /* extract diag(A) at creation (replaces lcg_smDcsr_get_diagonal, algebra_cuda.cu:40-57,
	                                 lcg_complex_cuda.cu:46-63) so lcgb200_jacobi_mx can be used */
};

/* Copies the CSR arrays (int32 row_ptr[n+1], int32 col[nnz], double|double2 val[nnz], base 0) to the current
 * device and builds the row tiles the SpMV kernel streams.  `location` says where the three arrays live. */
int lcgb200_csr_create(lcgb200_csr_t* out, int n, int nnz, const int* row_ptr, const int* col, const void* val,
	int value_type, int location, unsigned flags);
/* Rectangular variant for a row block of a partitioned matrix: n_rows local rows, columns index an extended
 * vector of n_cols >= n_rows entries (local entries first, then ghost entries). */
int lcgb200_csr_create_rect(lcgb200_csr_t* out, int n_rows, int n_cols, int nnz, const int* row_ptr, const int* col,
	const void* val, int value_type, int location, unsigned flags);
int lcgb200_csr_destroy(lcgb200_csr_t A);
/* user pointer handed to the progress callback as `instance` when the sentinel operator is used */
int lcgb200_csr_set_user(lcgb200_csr_t A, void* user_instance);
/* diag(A) to a HOST array of n values (double or interleaved complex) */
int lcgb200_csr_get_diagonal(lcgb200_csr_t A, void* diag_host);
/* the IC(0) factor of a handle created with LCGB200_CSR_IC0: number of entries of L, its CSR arrays on the HOST (any pointer
 * may be NULL; values double / interleaved complex in the handle's precision) and the number of dependency levels of the
 * forward (L) and backward (L^T) solves */
int lcgb200_csr_get_ic0(lcgb200_csr_t A, int* lnz, int* row_ptr_host, int* col_host, void* val_host, int* n_levels_lower, int* n_levels_upper);
/* z = (L L^T)^-1 r on device vectors: one application of the IC(0) preconditioner (two sparse triangular solves) */
int lcgb200_csr_ic0_apply(lcgb200_csr_t A, const void* r_dev, void* z_dev, void* stream);
/* the factorisation alone, on the host, in place on the lower triangle given as CSR (row_ptr[n+1], ascending columns, the
 * diagonal last in every row): the arithmetic of preconditioner.cpp:33-160 / preconditioner_cuda.cu:40-270 */
int lcgb200_ic0_factor_host(int n, const int* row_ptr, const int* col, void* val, int value_type);
/* y = op(A) x on device vectors; op: 0 = N, 1 = T, 2 = H (T/H need LCGB200_CSR_TRANSPOSE).  stream = cudaStream_t */
int lcgb200_csr_spmv(lcgb200_csr_t A, const void* x_dev, void* y_dev, int op, void* stream);
/* y = A x fused with the dot products the solvers take from it: dots[0] = w.y (w = x if w_dev is NULL),
 * dots[1] = y.y, dots[2] = x.y (real: plain sums; complex: conj-first inner products, 2 doubles each).
 * dots_dev receives 3 (real) or 6 (complex) doubles.  This is exactly the kernel the solvers launch. */
int lcgb200_csr_spmv_dot(lcgb200_csr_t A, const void* x_dev, void* y_dev, const void* w_dev, double* dots_dev, void* stream);
/* storage format the SpMV streams (LCGB200_CSR_COMPRESS): *compressed = 0 plain CSR (S+4 bytes per entry), 1 dictionary
 * codes (2 bytes per entry), 2 row patterns (1 byte per ROW: rows with identical (col - row, value) sequences share a
 * pattern; at most 253 patterns of at most 64 entries); the dictionary sizes; and the bytes one SpMV launch moves in that
 * format (matrix + row_ptr + x + y) */
int lcgb200_csr_format(lcgb200_csr_t A, int* compressed, int* n_values, int* n_offsets, long long* stream_bytes);
/* which kernel walks the row patterns (level 2): *kernel = 0 none, 1 the general chain kernel, 2 the box kernel (dense box
 * stencils on grids whose lines align with the threads' columns), 3 the plane-marching kernel (LCGB200_PAT_MARCH=1 in the
 * environment at creation); *stride = rows between the rows a thread owns (the line stride nx of a grid), *n_patterns = distinct rows */
int lcgb200_csr_pattern_kernel(lcgb200_csr_t A, int* kernel, int* stride, int* n_patterns);
/* bytes the SpMV kernel must move per launch by SURVEY.md §8(d): nnz*(S+4) + (n+1)*4 + 2*n*S */
long long lcgb200_csr_spmv_bytes(lcgb200_csr_t A);
int lcgb200_csr_info(lcgb200_csr_t A, int* n_rows, int* n_cols, int* nnz, int* n_tiles, int* lanes_per_row);

/* =====================================================================================================
 * Data step in front of the path (what every reference GPU sample does before solving, sample8.cu:30-64,169-173)
 * ===================================================================================================== */
/* replaces cusparseXcoo2csr (sample8.cu:169, sample9.cu:146): ascending row indices of nnz entries -> row_ptr[n+1], on the device */
int lcgb200_coo2csr(const int* rows_dev, int nnz, int n, int* row_ptr_dev, void* stream);
/* reads data/case_*_A (data/README:1-10; readers sample8.cu:30-52, sample9.cu:30-52): N, nz, the COO triplets (split, row-sorted)
 * and the right-hand side into malloc'ed HOST arrays (free with lcgb200_free_host); value_type LCGB200_REAL or LCGB200_COMPLEX */
int lcgb200_read_case(const char* path_A, int value_type, int* n_out, int* nz_out, int** rows_out, int** cols_out, void** vals_out, void** rhs_out);
void lcgb200_free_host(void* p);
/* the element-wise device helpers the reference's samples build their Jacobi Mx callbacks from (algebra_cuda.h:45-84,
 * lcg_complex_cuda.h:188-274; usage sample10.cu:117,193): c = a * b (op 0), c = a / b (op 1), c = conj(a) (op 2, b unused)
 * on device arrays of value_type LCGB200_REAL / LCGB200_COMPLEX / LCGB200_COMPLEX_FLOAT; A_diag[i] = A[i, i] of a device CSR
 * matrix (rows without a diagonal entry keep what diag_dev held); a = min(max(a, low), hig).  Asynchronous on `stream`. */
int lcgb200_vec_elementwise(int op, int value_type, const void* a_dev, const void* b_dev, void* c_dev, int n, void* stream);
int lcgb200_diagonal_of_csr(int value_type, const int* row_ptr_dev, const int* col_dev, const void* val_dev, int n, void* diag_dev, void* stream);
int lcgb200_set2box(const double* low_dev, const double* hig_dev, double* a_dev, int n, void* stream);
/* host COO triplets (row-sorted) -> operator handle; the COO -> CSR compression runs on the device */
int lcgb200_csr_create_from_coo(lcgb200_csr_t* out, int n, int nnz, const int* rows, const int* cols, const void* vals, int value_type, unsigned flags);

/* =====================================================================================================
 * Row-partitioned systems over several GPUs (new; the reference is single-device — SURVEY.md §8(e)).
 * One process per GPU.  Every rank creates its block with lcgb200_csr_create_rect (n_rows local rows; columns
 * index the extended vector [n_rows local entries | ghost entries grouped by owning peer]) and attaches the
 * exchange plan.  The solvers then run unchanged: before each SpMV the ghost entries are fetched from their
 * owners (ncclSend/ncclRecv over NVLink), and each fused reduction is summed over the ranks (ncclAllReduce on
 * 1-8 doubles) before its scalar epilogue.  m, B (low, hig) handed to lcgb200_solve are the rank's row slices.
 * ===================================================================================================== */
typedef struct lcgb200_comm_s* lcgb200_comm_t;
#define LCGB200_COMM_ID_BYTES 128
/* rank 0 draws the id (ncclGetUniqueId) and ships it to the other ranks by any means (MPI, torch.distributed, a file) */
int lcgb200_comm_unique_id(void* id_out, int capacity);
/* collective over all ranks; binds to the calling thread's current CUDA device */
int lcgb200_comm_create(lcgb200_comm_t* out, int rank, int size, const void* unique_id);
int lcgb200_comm_destroy(lcgb200_comm_t comm);
int lcgb200_comm_stats(lcgb200_comm_t comm, int* halo_exchanges, int* allreduces);
/* n_global: rows of the whole system (the abs_diff test divides by it, lcg.cpp:208).  For each of the n_peers
 * neighbours: its rank, how many of MY entries it needs (send_counts) — their local row indices concatenated in
 * send_idx (host array; a consecutive run is sent in place, anything else is packed by a gather kernel) — and how
 * many ghost entries I receive from it (recv_counts), stored in peer order behind the local entries.
 * sum(recv_counts) must equal n_cols - n_rows of the handle. */
int lcgb200_csr_set_partition(lcgb200_csr_t A, lcgb200_comm_t comm, long long n_global, int n_peers, const int* peer_ranks,
	const int* send_counts, const int* send_idx, const int* recv_counts);

/* Partitioned complex systems.  (1) clcg BiCG multiplies by A^H (clcg.cpp:188): build the rows [r0, r1) of A^T (values NOT
 * conjugated) as a second rectangular handle with its own communicator and partition, and attach it; A does not own At.
 * (2) The random shadow residual of complex CGS/BICGSTAB/TFQMR is ONE rand() sequence over the whole vector
 * (lcg_complex.cpp:118-127): tell the block where it starts so that every rank draws its slice of that sequence. */
int lcgb200_csr_attach_transpose(lcgb200_csr_t A, lcgb200_csr_t At);
int lcgb200_csr_set_row_offset(lcgb200_csr_t A, long long first_global_row);

/* NVLink peer-memory transport (optional, ranks on one NVLink/NVSwitch node): every rank exports the CUDA-IPC handle of
 * its communication window after lcgb200_csr_set_partition, the handles are exchanged by the caller (all-gather), and
 * lcgb200_comm_p2p_attach maps the peers' windows.  From then on real solves push halo entries and reduction totals
 * straight into the peers' memory from inside the kernels (no NCCL call on the iteration path).
 * handles: size x LCGB200_IPC_HANDLE_BYTES, in rank order; n_ghost_of_rank: size values (each rank's *n_ghost_out);
 * remote_off: for each of MY n_peers neighbours, in the order given to lcgb200_csr_set_partition, the offset at which
 * my entries start inside THAT rank's ghost region. */
#define LCGB200_IPC_HANDLE_BYTES 64
int lcgb200_comm_p2p_handle(lcgb200_comm_t comm, void* handle_out, long long* n_ghost_out);
int lcgb200_comm_p2p_attach(lcgb200_comm_t comm, const void* handles, const long long* n_ghost_of_rank, const long long* remote_off);
/* back to the NCCL transport: every rank must call it when the attach failed on ANY rank (the transports must agree) */
int lcgb200_comm_p2p_detach(lcgb200_comm_t comm);

/* Sentinel callbacks: never called; their ADDRESS selects the built-in operator.  instance = lcgb200_csr_t. */
void lcgb200_csr_ax(void* instance, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t x, lcgb200_dnvec_t Ax, const int n, const int nz);
void lcgb200_jacobi_mx(void* instance, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t x, lcgb200_dnvec_t Mx, const int n, const int nz);
void lcgb200_ic0_mx(void* instance, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t x, lcgb200_dnvec_t Mx, const int n, const int nz);   /* needs LCGB200_CSR_IC0 */
void lcgb200_ic0_cmx(void* instance, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t x, lcgb200_dnvec_t Mx, const int n, const int nz, int oper_t);
void lcgb200_csr_cax(void* instance, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t x, lcgb200_dnvec_t Ax, const int n, const int nz, int oper_t);
void lcgb200_jacobi_cmx(void* instance, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t x, lcgb200_dnvec_t Mx, const int n, const int nz, int oper_t);

/* =====================================================================================================
 * Reference-shaped entry points
 * ===================================================================================================== */
/* replaces lcg_solver_cuda (lcg_cuda.h:81-83, lcg_cuda.cu:40-58).  The reference only dispatches CG and CGS
 * here (anything else -> CG); we additionally accept BICGSTAB and BICGSTAB2 (CPU-only in the reference,
 * lcg.cpp:629-1034).  m, B: HOST arrays of n_size doubles; m is overwritten in place on every exit. */
int lcgb200_solver_cuda(lcgb200_axfunc_cuda_ptr Afp, lcgb200_progress_cuda_ptr Pfp, double* m, const double* B,
	const int n_size, const int nz_size, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);
/* replaces lcg_solver_preconditioned_cuda (lcg_cuda.h:104-106, lcg_cuda.cu:64-69); solver_id ignored as there */
int lcgb200_solver_preconditioned_cuda(lcgb200_axfunc_cuda_ptr Afp, lcgb200_axfunc_cuda_ptr Mfp, lcgb200_progress_cuda_ptr Pfp,
	double* m, const double* B, const int n_size, const int nz_size, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);
/* replaces lcg_solver_constrained_cuda (lcg_cuda.h:129-131, lcg_cuda.cu:76-81).  The reference's CUDA lpg is
 * broken; semantics follow the CPU lpg / lspg (lcg.cpp:1054-1447). */
int lcgb200_solver_constrained_cuda(lcgb200_axfunc_cuda_ptr Afp, lcgb200_progress_cuda_ptr Pfp, double* m, const double* B,
	const double* low, const double* hig, const int n_size, const int nz_size, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);
/* replaces clcg_solver_cuda (clcg_cuda.h:81-83, clcg_cuda.cu:42-60): BICG, BICG_SYM as there, plus CGS,
 * BICGSTAB, TFQMR (CPU-only in the reference, clcg.cpp:366-882).  m, B: HOST cuDoubleComplex arrays. */
int lcgb200_csolver_cuda(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_cprogress_cuda_ptr Pfp, void* m, const void* B,
	const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);
/* replaces clcg_solver_preconditioned_cuda (clcg_cuda.h:103-105, clcg_cuda.cu:70-84): CLCG_PCG only */
int lcgb200_csolver_preconditioned_cuda(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_caxfunc_cuda_ptr Mfp, lcgb200_cprogress_cuda_ptr Pfp,
	void* m, const void* B, const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);

/* replace the cuComplex overloads of clcg_solver_cuda / clcg_solver_preconditioned_cuda (clcg_cudaf.h:81-83, 103-105): BICG and
 * BICG_SYM / PCG in single-precision complex storage.  m, B: HOST arrays of n_size cuComplex.  Vectors and matrix values are
 * stored as floats (half the bytes per iteration); dot products, norms and the iteration scalars are carried in double.
 * The residual handed to the callback follows the reference's CUDA definition (clcg_cudaf.cu:142-176).  Built-in operator:
 * lcgb200_csr_cax / lcgb200_jacobi_cmx with an lcgb200_csr_t of value type LCGB200_COMPLEX_FLOAT as `instance`. */
typedef int (*lcgb200_cprogress_cudaf_ptr)(void* instance, const void* m_dev, const float converge,
	const lcgb200_cpara* param, const int n_size, const int nz_size, const int k);
int lcgb200_csolver_cudaf(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_cprogress_cudaf_ptr Pfp, void* m, const void* B,
	const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);
int lcgb200_csolver_preconditioned_cudaf(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_caxfunc_cuda_ptr Mfp, lcgb200_cprogress_cudaf_ptr Pfp,
	void* m, const void* B, const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub_handle, lcgb200_cusparse_t cus_handle, int solver_id);

/* ---- the reference's HOST-callback API (lcg.h:71-113, clcg.h:74-76) ----
 * Same arguments, dispatch and return codes as lcg_solver (CG, CGS, BICGSTAB, BICGSTAB2; anything else -> CGS, lcg.cpp:59-82),
 * lcg_solver_preconditioned (always PCG, lcg.cpp:87-91), lcg_solver_constrained (PG, SPG, lcg.cpp:121-140) and clcg_solver
 * (BICG, BICG_SYM, CGS, BICGSTAB, TFQMR; anything else -> CGS, clcg.cpp:46-74).  m, B (low, hig): HOST arrays.  Pass the
 * sentinels below with an lcgb200_csr_t as `instance` to run on the fused built-in operator; any other callback runs on
 * the HOST as in the reference (the vector is staged over PCIe for every call: correct, slow).  The progress callback
 * receives a HOST copy of the current solution. */
/* clcg_axfunc_ptr (clcg.h:40-41): layout 0 = MatNormal, 1 = MatTranspose; conjugate 0 = NonConjugate, 1 = Conjugate */
typedef void (*lcgb200_caxfunc_ptr)(void* instance, const void* x, void* prod_Ax, const int n_size, int layout, int conjugate);
typedef int (*lcgb200_cprogress_ptr)(void* instance, const void* m, const double converge, const lcgb200_cpara* param, const int n_size, const int k);
void lcgb200_csr_ax_host(void* instance, const double* x, double* prod_Ax, const int n_size);        /* sentinel */
void lcgb200_jacobi_mx_host(void* instance, const double* x, double* prod_Mx, const int n_size);     /* sentinel */
void lcgb200_ic0_mx_host(void* instance, const double* x, double* prod_Mx, const int n_size);        /* sentinel: built-in IC(0) */
void lcgb200_csr_cax_host(void* instance, const void* x, void* prod_Ax, const int n_size, int layout, int conjugate);   /* sentinel */
int lcgb200_solver(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const int n_size,
	const lcgb200_para* param, void* instance, int solver_id);
int lcgb200_solver_preconditioned(lcgb200_axfunc_ptr Afp, lcgb200_axfunc_ptr Mfp, lcgb200_progress_ptr Pfp, double* m, const double* B,
	const int n_size, const lcgb200_para* param, void* instance, int solver_id);
/* replace the stand-alone lcg() (lcg.h:135-137) and lcgs() (lcg.h:166-169): CG / CGS with optional caller-owned HOST work
 * vectors (any of them may be NULL).  The solve's work vectors live on the device; the caller's arrays receive their
 * final contents (lcg: Gk = gradient A m - B, Dk = direction, ADk = A Dk; lcgs: RK, R0T, PK, AX, UK, QK, WK). */
int lcgb200_lcg(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const int n_size,
	const lcgb200_para* param, void* instance, double* Gk, double* Dk, double* ADk);
int lcgb200_lcgs(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const int n_size,
	const lcgb200_para* param, void* instance, double* RK, double* R0T, double* PK, double* AX, double* UK, double* QK, double* WK);
int lcgb200_solver_constrained(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const double* low, const double* hig,
	const int n_size, const lcgb200_para* param, void* instance, int solver_id);
int lcgb200_csolver(lcgb200_caxfunc_ptr Afp, lcgb200_cprogress_ptr Pfp, void* m, const void* B, const int n_size,
	const lcgb200_cpara* param, void* instance, int solver_id);

/* =====================================================================================================
 * Handle-shaped entry points (same engine; vectors may already be on the device)
 * A handle caches the workspace, the device state block and the pinned read-back buffers of its solves: ONE solve at a
 * time per handle (use one handle per concurrent solve; the matrix arrays can be shared by creating it from device arrays).
 * ===================================================================================================== */
typedef struct lcgb200_info {
	int iterations;        /* k handed to the last convergence check (what the reference reports through Pfp) */
	int checks;            /* number of loop-head convergence checks performed */
	int spmv_launches;     /* SpMV kernels launched (including skipped no-op launches after convergence) */
	int kernel_launches;   /* all kernels of ours launched by this call */
	double residual;       /* residual at the last check (the `converge` value of the callbacks) */
	double device_ms;      /* CUDA-event time from the first to the last kernel of the solve */
	double total_ms;       /* host wall time of the whole call, copies included */
	/* filled when lcgb200_set_profile(1): CUDA-event time summed over the SpMV(+fused dot) launches and over the
	 * fused vector kernels of this solve, and how many launches of each were timed */
	double spmv_ms, vec_ms;
	int spmv_timed, vec_timed;
} lcgb200_info;

enum {
	LCGB200_VEC_DEVICE = 1,   /* m, B (low, hig) are device pointers */
	LCGB200_USE_JACOBI = 2,   /* PCG: built-in Jacobi z = r/diag (needs LCGB200_CSR_JACOBI) */
	LCGB200_USE_IC0 = 4       /* PCG: built-in IC(0) z = L^-T L^-1 r (needs LCGB200_CSR_IC0) */
};

/* real solvers on the built-in operator: solver_id in LCGB200_CG..LCGB200_SPG (low/hig only for PG, SPG) */
int lcgb200_solve(lcgb200_csr_t A, int solver_id, double* m, const double* B, const double* low, const double* hig,
	const lcgb200_para* param, lcgb200_progress_cuda_ptr Pfp, unsigned flags, void* stream, lcgb200_info* info);
/* complex solvers on the built-in operator: solver_id in LCGB200_CBICG..LCGB200_CPCG; m, B in the handle's precision
 * (cuDoubleComplex for LCGB200_COMPLEX, cuComplex for LCGB200_COMPLEX_FLOAT) */
int lcgb200_csolve(lcgb200_csr_t A, int solver_id, void* m, const void* B, const lcgb200_cpara* param,
	lcgb200_cprogress_cuda_ptr Pfp, unsigned flags, void* stream, lcgb200_info* info);

/* =====================================================================================================
 * Settings / diagnostics
 * ===================================================================================================== */
/* seed used for the random shadow residual of complex CGS/BICGSTAB/TFQMR; 0 = time(0) as the reference does
 * (lcg_complex.cpp:118-127).  The sequence is libc srand()/rand(), drawn on the host, like the reference. */
void lcgb200_set_shadow_seed(long seed);
/* residual definition of the complex solvers: 0 (default) = reference CPU (|r|^4/max(|m|^4,1), clcg.cpp:112-147),
 * 1 = reference CUDA (|r|^2/max(|m|,1)^2, clcg_cuda.cu:145-176) */
void lcgb200_set_complex_residual_mode(int mode);
/* iterations enqueued between two host polls of the device convergence flag when no progress callback is set */
void lcgb200_set_poll_interval(int iterations);
/* 1 (default): systems small enough to stay cache-resident (<= 65536 rows, <= 2M non-zeros; CG, Jacobi-PCG and the complex
 * BICG_SYM / Jacobi-PCG) run several whole iterations per cooperative launch; 0: always the streaming kernels */
void lcgb200_set_fused_small(int on);
/* Cross-GPU waits of the NVLink peer-memory transport (a rank waiting for a neighbour's halo or reduction totals) give up
 * after this many milliseconds (default: environment LCGB200_SPIN_TIMEOUT_MS, else 30000; 0 = wait for ever); the limit is
 * 20x longer while a progress callback or a host-side operator keeps the host inside every iteration.  A timeout ends the
 * solve with LCGB200_UNKNOWN_ERROR and a message (lcgb200_last_error), and POISONS the communicator: its sequence counters
 * may no longer agree with the peers', so later solves on it are refused — destroy it and create a new one. */
void lcgb200_set_spin_timeout_ms(long long ms);
/* CUDA graphs for batches of iterations when no host-visible sync point lies inside an iteration: -1 (default) =
 * environment LCGB200_GRAPHS or automatic (on for systems whose iteration is short enough for launch gaps to matter),
 * 0 = off, 1 = on */
void lcgb200_set_graphs(int mode);
/* programmatic dependent launch between the kernels of an iteration (the next kernel is scheduled while the current one
 * finishes its reduction and parks in griddepcontrol.wait): -1 (default) = environment LCGB200_PDL or on, 0 = off, 1 = on */
void lcgb200_set_pdl(int mode);
/* Persisting-L2 window over the iteration's work vectors for the duration of a solve (systems whose vectors fit the 126 MB
 * L2 while the matrix does not: 10^6-10^7 rows per GPU): -1 (default) = environment LCGB200_L2_PERSIST or automatic,
 * 0 = off, 1 = always.  Uses the device-wide persisting carve-out (cudaLimitPersistingL2CacheSize), released at the end
 * of the solve. */
void lcgb200_set_l2_persist(int mode);
/* Reference-order arithmetic (verification mode; single GPU, double precision): 1 = every later solve runs the second
 * build of the iteration loops, which reproduces the reference's x86-64 CPU arithmetic operation for operation — no fused
 * multiply-adds, SpMV row sums and dot products added left to right in index order (algebra.cpp:154-163,
 * lcg_complex.cpp:143-167) — so that iterates, residual histories and iteration counts are BIT-IDENTICAL to
 * lcg_solver / clcg_solver driven by a plain CSR callback.  O(n) dependent additions per dot product: meant for the
 * reference's sample systems (10^3-10^4 rows), where it turns "statistically indistinguishable" convergence counts of the
 * erratic recurrences (BiCG, BiCGSTAB, CGS) into equalities.  -1 (default) = environment LCGB200_REFERENCE_ORDER or off. */
void lcgb200_set_reference_order(int mode);
/* 1: bracket every kernel launch of a solve with CUDA events (per-kernel durations in lcgb200_info); costs a
 * little throughput, so bench.py uses it only for its roofline pass */
void lcgb200_set_profile(int on);
const char* lcgb200_last_error(void);
int lcgb200_version(void);
/* device helpers used by tests / bench: fill CSR rows [row0,row1) of a g^3 stencil directly on the device.
 * kind: 0 = 7pt Poisson, 1 = 27pt Poisson, 2 = 7pt convection-diffusion.  Pass NULL arrays
 * to only get the nnz count of the row range in *nnz_out. */
int lcgb200_gen_stencil(int kind, int g, long long row0, long long row1, int* row_ptr_dev, int* col_dev, double* val_dev,
	long long col_offset, long long* nnz_out, void* stream);
/* b[row0..row1) = (A x*)[row0..row1) with x*[i] = uint32(i*2654435761)/2^32, summed left to right per row */
int lcgb200_gen_rhs(int kind, int g, long long row0, long long row1, double* b_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCGB200_H */
