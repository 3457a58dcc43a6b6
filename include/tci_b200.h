/*
 * tci_b200.h -- C ABI of libtci_b200.so, the B200 (sm_100a) drop-in for the
 * TensorCI2 two-site update hot path of TensorCrossInterpolation.jl v0.9.19.
 *
 * The reference is pure Julia and has no FFI; its extension points are Julia
 * dispatch.  Each entry point below names the reference interface it stands in
 * for (file:line under /root/reference/src); INTEGRATION.md shows the Julia
 * `ccall` shim that binds them.
 *
 * Conventions (Julia's): matrices are column-major Float64; sizes are int64_t;
 * multi-indices and permutations are 1-based; index sets are passed flattened,
 * (len x count) column-major, i.e. multi-index q occupies I[len*q .. len*q+len).
 * Host buffers belong to the caller and are only used during the call.
 * Library objects are opaque handles with explicit destroy functions.
 * Every function returns TCI_OK (0) or a tci_status; tci_last_error(ctx) gives
 * the message (texts follow the reference's exceptions where one exists).
 * A tci_ctx drives one GPU or, created with ngpu > 1, several GPUs of one node from ONE process (the reference's
 * caller is a single Julia process): device_ids[0] owns the per-bond rrLU, and the stages that shard -- Pi
 * evaluation, the random-start global search, the row blocks of an MPO x MPO contraction -- are split over all of
 * them inside the library (peer stores over NVLink, NCCL for the gathers).  A context is used by one caller at a time.
 */
#ifndef TCI_B200_H
#define TCI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tci_ctx tci_ctx;
typedef struct tci_dmat tci_dmat; /* device matrix (m x n, leading dimension ld) */
typedef struct tci_lu tci_lu;     /* device-resident rrLU factorisation          */

typedef enum {
    TCI_OK = 0,
    TCI_ERR_CUDA = 1,        /* CUDA runtime failure                              */
    TCI_ERR_ARG = 2,         /* invalid argument                                  */
    TCI_ERR_NAN_L = 3,       /* "lu.L contains NaNs"  matrixlu.jl:164-166         */
    TCI_ERR_NAN_U = 4,       /* "lu.U contains NaNs"  matrixlu.jl:167-169         */
    TCI_ERR_CENTRE = 5,      /* "Invalid number of central indices" tensorci2.jl:307 */
    TCI_ERR_NO_DEVICE = 6,   /* no CUDA device: the library has no CPU fallback   */
    TCI_ERR_BUSY = 7,        /* context entered concurrently                      */
    TCI_ERR_UNSUPPORTED = 8,
    TCI_ERR_SINGULAR = 9     /* "Pivot matrix at bond b is singular!" (the `\` of tensorci2.jl:391 would throw) */
} tci_status;

int tci_version(void);

/* ---- context ------------------------------------------------------------ */
/* ngpu >= 1 GPUs of this node, device_ids[0] = owner (NULL: devices 0 .. ngpu-1).  ngpu > 1 needs peer access between
 * all of them (NVLink / NVSwitch) and NCCL (libnccl.so.2, bound at run time); TCI_ERR_UNSUPPORTED otherwise.          */
int tci_ctx_create(int ngpu, const int *device_ids, tci_ctx **out);
/* Handles (tci_dmat, tci_lu) keep the context alive: destroying it first is allowed, its resources are released with
 * the last handle (Julia finalizers run in arbitrary order).                                                          */
void tci_ctx_destroy(tci_ctx *ctx);
int tci_ctx_ngpu(tci_ctx *ctx);
const char *tci_last_error(tci_ctx *ctx); /* ctx may be NULL: last create error */
/* number of kernels this context launched since creation (bench gpu_launches) */
int64_t tci_ctx_launches(tci_ctx *ctx);
/* the cudaStream_t all work of this context is issued on (for event timing by the host layer) */
void *tci_ctx_stream(tci_ctx *ctx);
/* accumulated CUDA-event time per stage in ms; stages: 0 pi_eval, 1 rrlu, 2 luci,
 * 3 tt/mpo environments, 4 globalsearch, 5 gemm, 6 h2d, 7 d2h, 8 the rrLU kernel alone.  Stands in for the
 * time_ns() pairs around "Computing Pi"/"LU" (tensorci2.jl:530-550).              */
int tci_timers(tci_ctx *ctx, double *out, int64_t n, int reset);
/* kernels launched / stage times of member k of a multi-GPU context (k = 0: the owner) */
int64_t tci_ctx_member_launches(tci_ctx *ctx, int k);

/* ---- device matrices ---------------------------------------------------- */
int tci_dmat_create(tci_ctx *ctx, int64_t m, int64_t n, const double *host /* nullable */, tci_dmat **out);
/* As tci_dmat_create with a host matrix, but the upload runs on the context's copy stream and the call returns
 * once it is enqueued, so that it overlaps with a factorisation in flight (double buffering of the Julia-side
 * `Matrix{Float64}` handed to rrlu, matrixlu.jl:217-225).  `host` must stay valid -- and should be page-locked --
 * until the matrix is first consumed (tci_rrlu, tci_lu_rdiv, tci_dmat_fetch, tci_dmat_destroy order themselves
 * after the copy).                                                                                         */
int tci_dmat_create_async(tci_ctx *ctx, int64_t m, int64_t n, const double *host, tci_dmat **out);
int tci_dmat_shape(tci_dmat *a, int64_t *m, int64_t *n, int64_t *ld);
void *tci_dmat_ptr(tci_dmat *a); /* raw device pointer, for collectives on the host layer */
int tci_dmat_fetch(tci_dmat *a, double *host /* m x n, tight */);
int tci_dmat_destroy(tci_dmat *a);
/* reshape(A, m2, n2) of the column-major m x n matrix (m*n == m2*n2) as a new device matrix: the backward half-sweep
 * of sweep1site! folds the same T tensor as |I| x (d*|J|) instead of (|I|*d) x |J| (tensorci2.jl:417-428).           */
int tci_dmat_refold(tci_dmat *a, int64_t m2, int64_t n2, tci_dmat **out);
/* shrink the logical column count (n <= allocated columns); used after an all-gather of padded blocks */
int tci_dmat_resize_cols(tci_dmat *a, int64_t n);

/* non-owning device matrix on caller-managed memory (ld >= m, ld % 2 == 0, 16-byte aligned base) */
int tci_dmat_wrap(tci_ctx *ctx, void *dptr, int64_t m, int64_t n, int64_t ld, tci_dmat **out);

/* ---- targets: the function f being interpolated ------------------------- */
/* Replaces the Julia closure / BatchEvaluator object `f` (cachedtensortrain.jl:1,
 * batcheval.jl:67-83, docs/src/index.md:174-241).  kind_id: tci_targets.h.       */
int tci_target_builtin(tci_ctx *ctx, int kind_id, const double *params, int64_t nparams, const int64_t *localdims,
                       int64_t nsites, int64_t *target_id);
/* A user-defined target: the Julia closure `f` restated as CUDA source that defines
 *     __device__ double tci_user_f(const long long *x, int n, const double *params);
 * (x[0..n): 1-based local indices), compiled at run time with NVRTC for sm_100a (--fmad=false: products and sums round
 * separately, as in the closure) and evaluated by the same entry points as a built-in target.  nsites <= 128.
 * TCI_ERR_ARG with the compiler log if the source does not compile; TCI_ERR_UNSUPPORTED without libnvrtc.so.12.      */
int tci_target_source(tci_ctx *ctx, const char *source, const double *params, int64_t nparams, const int64_t *localdims,
                      int64_t nsites, int64_t *target_id);
/* TTCache(tt) (cachedtensortrain.jl:9-30): cores[s] is (dims3[3s], dims3[3s+1], dims3[3s+2]). */
int tci_tt_create(tci_ctx *ctx, int64_t nsites, const int64_t *dims3, const double *const *cores,
                  int64_t *target_id);
/* Core `site` (0-based) of a device-resident tensor train -- tci_tt_create, or the handle tci_fill_sitetensors returns:
 * dims3 (nullable) receives (Dl, d, Dr), out (nullable) the Dl*d*Dr values.  tci.sitetensors[site] read lazily: inside
 * optimize! the site tensors are only consumed by the global pivot finder (tensorci2.jl:788-800), on the device.       */
int tci_tt_fetch_core(tci_ctx *ctx, int64_t tt_id, int64_t site, int64_t *dims3, double *out);
/* Contraction(a, b) (contraction.jl:35-62): A[s] is (Da,s1,s2,Da'), B[s] is (Db,s2,s3,Db'). */
int tci_mpo_pair_create(tci_ctx *ctx, int64_t nsites, const int64_t *dimsA4, const double *const *A,
                        const int64_t *dimsB4, const double *const *B, int64_t *target_id);
/* CachedFunction{Float64,UInt128}(f, localdims) (cachedfunction.jl:8-63) as a device-resident memo: a target that wraps
 * target `inner_id` (any real-valued target of this context; it must outlive the wrapper) with a hash table of
 * 2^capacity_log2 slots (32 B each) in HBM, keyed like the reference by key(x) = sum_n coeffs[n] (x_n - 1), coeffs[n] =
 * prod_{m<n} localdims[m], as UInt128 ("Overflow in CachedFunction..." -> TCI_ERR_ARG, :22-24).  Every entry point that
 * takes a target accepts it; an evaluation looks all requested elements up, evaluates the wrapped target on the missing
 * points only, in one batch, and stores the new values (_batcheval_imp_for_batchevaluator, :117-171).  f must be pure.
 * Single-GPU contexts only (the memo lives in one GPU's HBM).                                                        */
int tci_target_cached(tci_ctx *ctx, int64_t inner_id, int capacity_log2, int64_t *target_id);
/* out[0] = length(cf.cache), out[1] = lookups answered from the memo, out[2] = lookups that evaluated f, out[3] = values
 * that found no free slot within the probe limit (returned correctly, not memoised).                                 */
int tci_target_cache_stats(tci_ctx *ctx, int64_t target_id, int64_t *out /* 4 */);
int tci_target_destroy(tci_ctx *ctx, int64_t target_id);
/* The elementwise function of a Contraction (`f` of contraction.jl:5-62, applied at :203-205 and :330-332).  A
 * device kernel cannot call a Julia closure, so functions are registered by id like the targets themselves:
 * TCI_F_NONE, TCI_F_AFFINE (a*x + b, rounded multiply then rounded add), TCI_F_ABS (|x|), TCI_F_SQUARE (x*x).
 * Only MPO-pair targets accept one (TCI_ERR_ARG otherwise).                                               */
enum { TCI_F_NONE = 0, TCI_F_AFFINE = 1, TCI_F_ABS = 2, TCI_F_SQUARE = 3 };
int tci_target_set_elementwise(tci_ctx *ctx, int64_t target_id, int kind, double a, double b);
/* f(x) for `count` full multi-indices (nsites x count): the scalar call
 * (bf::BatchEvaluatorAdapter)(indexset) batcheval.jl:11-13, TTCache/Contraction
 * evaluate cachedtensortrain.jl:130-146, contraction.jl:189-207.                  */
int tci_target_eval(tci_ctx *ctx, int64_t target_id, const int64_t *idx, int64_t count, double *out);

/* ---- (a) batched Pi / T evaluation --------------------------------------- */
/* filltensor / _batchevaluate_dispatch (tensorci2.jl:290-312, batcheval.jl:32-83)
 * and batchevaluate of TTCache / Contraction (cachedtensortrain.jl:151-215,
 * contraction.jl:236-335).  out[i, c, j] = f(I_i ++ c ++ J_j); M centre sites
 * nl+1..nl+M are expanded on the device, first centre index fastest; layout is
 * column-major (nI, d_{nl+1}, ..., d_{nl+M}, nJ).  nI*nJ == 0 is not an error
 * (empty result, batcheval.jl:40-42).  *maxabs (nullable) receives
 * max(|out|), NaN-propagating, i.e. maxabs(0, out) of util.jl:1-10.
 * out_host (nullable) receives the tight array; out_dev (nullable) receives a
 * device matrix of shape (nI*C) x nJ that can be handed to tci_rrlu.             */
int tci_pi_eval(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                int64_t nr, int64_t nJ, int64_t M, double *out_host, tci_dmat **out_dev, double *maxabs);

/* Same evaluation written into columns [col0, col0+nJ) of an existing device matrix whose row count
 * is nI*C: the column-block form used when Pi is sharded over GPUs; every rank fills its
 * own block of the buffer that is all-gathered for the rrLU owner.                                  */
int tci_pi_eval_into(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                     int64_t nr, int64_t nJ, int64_t M, tci_dmat *dst, int64_t col0, double *maxabs);

/* Environments of a TT / MPO-pair target as device matrices of their own, so that the two chains of a
 * contraction Pi can be sharded independently over GPUs ("MPO x MPO contraction: row blocks"):
 * side 0 = evaluateleft over the first `len` sites, side 1 = evaluateright over the last `len` sites
 * (cachedtensortrain.jl:77-128, contraction.jl:112-176).  idx is (len x count), entry q contiguous.  Column
 * col0+q of dst receives the environment of entry q; dst->m must equal tci_env_dim (the bond dimension of a
 * TT, Da*Db of an MPO pair, 1 for len == 0).  Analytic targets: TCI_ERR_ARG.                          */
int tci_env_dim(tci_ctx *ctx, int64_t target_id, int side, int64_t len, int64_t *D);
int tci_env_eval(tci_ctx *ctx, int64_t target_id, int side, const int64_t *idx, int64_t len, int64_t count,
                 tci_dmat *dst, int64_t col0);
/* The M = 0 Pi from such environments (cachedtensortrain.jl:211-212, contraction.jl:328):
 * dst[:, col0 + j] = left[:, l0 + i]^T right[:, r0 + j], i < nI, j < nJ; dst->m must be nI.  *maxabs as in
 * tci_pi_eval (nullable).                                                                              */
int tci_pi_from_envs(tci_ctx *ctx, tci_dmat *left, int64_t l0, int64_t nI, tci_dmat *right, int64_t r0, int64_t nJ,
                     tci_dmat *dst, int64_t col0, double *maxabs);

/* ---- (b) rank-revealing LU / MatrixLUCI ---------------------------------- */
/* rrlu(A; maxrank, reltol, abstol, leftorthogonal) (matrixlu.jl:194-225) with the
 * full-pivot search and Schur updates of matrixlu.jl:1-32,98-181.  Exactly one of
 * A_host / A_dev is non-NULL; A_dev is factorised IN PLACE and owned by *factors
 * afterwards (do not destroy it separately when factors != NULL).
 * maxrank <= 0 or > min(m,n) means min(m,n).  exact_mode = 1 reproduces the
 * reference arithmetic (rounded multiply, rounded subtract, true division);
 * exact_mode = 0 uses fused multiply-add.
 * Outputs: rowperm[m], colperm[n] (1-based, = lu.rowpermutation/colpermutation),
 * *npivot, *error (= lu.error), pivoterrors[npivot+1] (caller provides
 * min(m,n)+1 doubles; = pivoterrors(lu) matrixlu.jl:394-416).                     */
int tci_rrlu(tci_ctx *ctx, const double *A_host, tci_dmat *A_dev, int64_t m, int64_t n, int64_t maxrank,
             double reltol, double abstol, int leftorthogonal, int exact_mode, int64_t *rowperm, int64_t *colperm,
             int64_t *npivot, double *error, double *pivoterrors, tci_lu **factors /* nullable */);
/* lu.L (m x r) and lu.U (r x n), i.e. left/right(lu; permute=false) matrixlu.jl:374-392 */
int tci_lu_fetch(tci_lu *lu, double *L /* nullable */, double *U /* nullable */);
/* left(luci) (m x r) / right(luci) (r x n)  matrixluci.jl:40-84 */
int tci_luci_left(tci_lu *lu, double *out_host /* nullable */, tci_dmat **out_dev /* nullable */);
int tci_luci_right(tci_lu *lu, double *out_host /* nullable */, tci_dmat **out_dev /* nullable */);
/* B * A^-1 for a square A factorised to full rank by tci_rrlu (npivot == m == n): the solve of
 * setsitetensor!, T = Pi1 * P^-1, `transpose(transpose(P) \\ transpose(Pi1))` at tensorci2.jl:391.
 * The reference leaves it to LAPACK gesv (partial pivoting, version unpinned -- SURVEY 8c); here the
 * full-pivot factors are reused: X[:, rowperm] = B[:, colperm] U^-1 L^-1.  B (rows x m) is not modified. */
int tci_lu_rdiv(tci_lu *lu, tci_dmat *B, double *out_host /* nullable */, tci_dmat **out_dev /* nullable */);
/* Completion of a rook-search factorisation (arrlu, matrixlu.jl:274-288) from the factors of its r x r pivot block:
 * L2 = A21 U11^-1 (cols2Lmatrix!, :314-335; A21 is (m2 x r)) and U2 = L11^-1 A12 (rows2Umatrix!, :337-358; A12 is
 * (r x n2)); either may be NULL.  Results are written tight to L2_host (m2 x r) / U2_host (r x n2).               */
int tci_lu_complete(tci_lu *lu, tci_dmat *A21, tci_dmat *A12, double *L2_host, double *U2_host);
int tci_lu_destroy(tci_lu *lu);

/* ---- fused entry points of the driver's inner loop ------------------------- */
/* The `:full` branch of updatepivots! (tensorci2.jl:529-551) in one call: Pi = f(Icombined x Jcombined) is evaluated
 * into HBM (sharded over the context's GPUs when the cost model says it pays), factorised in place by the rrLU on the
 * owner, and only the permutations / pivot errors / max|Pi| come back -- one host synchronisation per bond.  I, J are
 * the already expanded index sets (kronecker + union, :526-527).  Outputs as tci_rrlu; rowindices = rowperm[0:npivot],
 * colindices = colperm[0:npivot]; *maxabs = max|Pi| for updatemaxsample! (:538).  *factors (nullable) owns Pi.        */
int tci_bond_update(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                    int64_t nr, int64_t nJ, int64_t maxrank, double reltol, double abstol, int leftorthogonal,
                    int exact_mode, int64_t *rowperm, int64_t *colperm, int64_t *npivot, double *error,
                    double *pivoterrors, double *maxabs, tci_lu **factors);
/* The bond loop of one half-sweep of sweep2site! (tensorci2.jl:866-907): updatepivots! (:510-607, `:full` search) for
 * b = 1 .. n-1 (forward != 0, leftorthogonal factorisations) or n-1 .. 1 (backward) in ONE call.  Per bond:
 * Icombined = union(kronecker(Iset[b], d_b), extraI[b+1]), Jcombined = union(kronecker(d_{b+1}, Jset[b+1]), extraJ[b])
 * (:526-527; extra sets nullable = strictlynested), Pi-evaluation -> rrLU with maxrank = maxbonddim (<= 0: none),
 * Iset[b+1] / Jset[b] replaced by the selected rows / columns (:597-598), updateerrors (:161-169).  Index sets are
 * ragged arrays: Iset[b] is (b x nI[b]), Jset[b] is ((n-1-b) x nJ[b]), multi-index contiguous, b = 0 .. n-1 (0-based).
 * The call returns the new set sizes and the length of the pivot-error vector; tci_sweep2site_fetch then copies the
 * sets into caller-allocated arrays of those sizes, with bonderrors (n-1), pivoterrors, max|Pi| over the bonds
 * (updatemaxsample!, :538) and, per bond in the order visited, (bond, rows, columns, npivot) (4 (n-1) values).
 * Not formed: the two site tensors updatepivots! sets when there are no extra sets (:599-602) -- the next bond
 * invalidates them and fillsitetensors! (:909-911) rebuilds all of them at the end of sweep2site!.                    */
int tci_sweep2site_half(tci_ctx *ctx, int64_t target_id, int forward, const int64_t *const *Iset, const int64_t *nI,
                        const int64_t *const *Jset, const int64_t *nJ, const int64_t *const *extraI,
                        const int64_t *nextraI, const int64_t *const *extraJ, const int64_t *nextraJ, double reltol,
                        double abstol, int64_t maxbonddim, int exact_mode, int64_t *nI_out, int64_t *nJ_out,
                        int64_t *npivoterrors);
int tci_sweep2site_fetch(tci_ctx *ctx, int64_t *const *Iout, int64_t *const *Jout, double *bonderrors,
                         double *pivoterrors, double *maxsample, int64_t *trace /* nullable */);
/* fillsitetensors! (globalsearch.jl:97-103) = setsitetensor!(tci, f, b) for every site (tensorci2.jl:367-394):
 * T_b = Pi1_b P_b^-1 with Pi1_b = f(Iset[b] x sigma_b x Jset[b]) and P_b = f(Iset[b+1] x Jset[b]); the last tensor is
 * Pi1 itself.  All evaluations, full-rank factorisations and solves are queued back to back and synchronised once.
 * Iset[b]: (b x nI[b]), Jset[b]: ((nsites-1-b) x nJ[b]); T_out[b] (nullable entries / nullable array) receives
 * nI[b]*d_b*nJ[b] doubles; *maxabs = max over all |Pi1| (updatemaxsample!, :375).  *tt_id (nullable) receives the
 * result as a device-resident tensor-train target (the current_tt of the global pivot finder); destroy it with
 * tci_target_destroy.  Errors: "Pivot matrix at bond b is not square!" (TCI_ERR_ARG, :388), TCI_ERR_SINGULAR.       */
int tci_fill_sitetensors(tci_ctx *ctx, int64_t target_id, int64_t nsites, const int64_t *const *Iset,
                         const int64_t *nI, const int64_t *const *Jset, const int64_t *nJ, double *const *T_out,
                         double *maxabs, int64_t *tt_id);

/* ---- dense FP64 GEMM on device matrices (building block of (b),(c)) ------- */
/* C = alpha * op(A) * op(B) + beta * C ; host arrays, column-major, tight.  Exposed
 * for benchmarks and tests of the kernel that replaces OpenBLAS dgemm in
 * matrixluci.jl:41,45, cachedtensortrain.jl:207,212 and contraction.jl:92.        */
int tci_dgemm_host(tci_ctx *ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                   const double *A, const double *B, double beta, double *C);

/* FP64 roofline denominators measured on this GPU: out[0] = DFMA pipe TFLOP/s (register-resident FMA loop),
 * out[1] = FP64 tensor path TFLOP/s (register-resident mma.sync.m8n8k4.f64 loop; tcgen05 has no FP64 kind).       */
int tci_fp64_peak(tci_ctx *ctx, double *out /* 2 */);

/* ---- (c) contraction ------------------------------------------------------ */
/* One zip-up step (contraction.jl:455-464): R (chi,Da,Db), A (Da,s1,s2,Da'),
 * B (Db,s2,s3,Db') -> C as the (chi*s1*s3) x (Da'*Db') matrix that is factorised next. */
int tci_contract_zipup_site(tci_ctx *ctx, const double *R, int64_t chi, int64_t Da, int64_t Db, const double *A,
                            int64_t s1, int64_t s2, int64_t Dan, const double *B, int64_t s3, int64_t Dbn,
                            double *C_host /* nullable */, tci_dmat **C_dev /* nullable */);
/* _contractsitetensors (contraction.jl:338-349): out (Da*Db, s1, s3, Da'*Db') */
int tci_contract_naive_site(tci_ctx *ctx, const double *A, int64_t Da, int64_t s1, int64_t s2, int64_t Dan,
                            const double *B, int64_t Db, int64_t s3, int64_t Dbn, double *out_host);

/* ---- global pivot search --------------------------------------------------- */
/* DefaultGlobalPivotFinder call (globalpivotfinder.jl:143-195) with the start
 * points drawn by the caller's rng (n x nsearch): for every start the star of
 * sum_p d_p probes |f(x) - tt(x)|, first maximum kept, accepted if > threshold
 * (= abstol * tolmarginglobalsearch), truncated to the first maxn in start order.
 * tt_id: the current tensor train as a device-resident handle (tci_tt_create, or the one tci_fill_sitetensors
 * returns) -- nothing is uploaded per call.  The per-start arg-max runs on the device; only nsearch (error, index)
 * records come back.  On a multi-GPU context the starts are split into contiguous blocks over the GPUs and the records
 * gathered by one ncclAllGather.
 * mode: 1 = every probe through the ordered left-to-right chain (bit-identical to evaluate(tt, x),
 * abstracttensortrain.jl:124-132); 2 = prefix / suffix environments of the start points + one GEMM per site (~n times
 * fewer flops; values agree with mode 1 to rounding, pivots identical away from near-ties); 0 = mode 1 below 32768
 * probes (the reference's default nsearch = 5), mode 2 above.
 * pivots_out: n x maxn, errs_out: maxn, start_idx_out (nullable): maxn, the 0-based start each accepted pivot came from. */
int tci_globalsearch(tci_ctx *ctx, int64_t target_id, int64_t tt_id, const int64_t *starts, int64_t nsearch,
                     double threshold, int64_t maxn, int mode, int64_t *pivots_out, double *errs_out,
                     int64_t *start_idx_out, int64_t *nfound);
/* The same call with the start points drawn inside the library by the injected counter-based generator that replaces
 * Julia's rng at globalpivotfinder.jl:156 (start s, site p = 1 + floor(tci_uniform01(seed, (call*1000003 + s)*1009 + p)
 * * d_p), include/tci_targets.h; `call` counts the finder calls of a run, from 1): every GPU draws its own block of
 * starts on the device, so nothing is generated or uploaded by the host.  Results as tci_globalsearch.               */
int tci_globalsearch_counter(tci_ctx *ctx, int64_t target_id, int64_t tt_id, uint64_t seed, uint64_t call,
                             int64_t nsearch, double threshold, int64_t maxn, int mode, int64_t *pivots_out,
                             double *errs_out, int64_t *start_idx_out, int64_t *nfound);
/* Host-only pieces of the sharded stages, exported so that the partitioning and the selection can be tested without a
 * GPU: the block [lo, hi) of `rank` when n items are split over `world` ranks in contiguous blocks whose starts are
 * multiples of `align`; and the selection of :180-188 replayed on gathered (error, probe index) records.             */
int tci_shard_range(int64_t n, int world, int rank, int64_t align, int64_t *lo, int64_t *hi);
/* The order in which a sharded TT / contraction Pi deals rows (side 0) or columns (side 1) to the GPUs: environments are
 * shared by entries with a common partial index (the Dict memo of contraction.jl:112-176, cachedtensortrain.jl:77-128),
 * so rows are taken in lexicographic order of their multi-indices (first site most significant), columns with the LAST
 * site most significant, and every GPU gets a contiguous block (tci_shard_range) of that order.  idx: (len x count);
 * perm[q] = caller's index of the q-th entry of the order (stable).  Host only.                                        */
int tci_shard_order(const int64_t *idx, int64_t len, int64_t count, int side, int64_t *perm);
int tci_globalsearch_select(const double *rec_err, const int64_t *rec_idx, int64_t nsearch, const int64_t *starts,
                            int64_t nsites, const int64_t *localdims, double threshold, int64_t maxn,
                            int64_t *pivots_out, double *errs_out, int64_t *start_idx_out, int64_t *nfound);
/* evaluate(tt, x) for `count` points, left-to-right (abstracttensortrain.jl:124-132) */
int tci_tt_evaluate(tci_ctx *ctx, int64_t nsites, const int64_t *dims3, const double *const *cores,
                    const int64_t *idx, int64_t count, double *out);

/* ---- ComplexF64 value type -------------------------------------- */
/* The reference is generic in the value type and its contraction / conversion tests run on ComplexF64
 * (test_contraction.jl:39-46, test_matrixlu.jl:39-52).  A Matrix{ComplexF64} crosses the ABI as Julia stores it:
 * interleaved (re, im) pairs, column-major; on the device a complex m x n matrix is a tci_dmat of 2m x n doubles
 * (tci_dmat_create(ctx, 2*m, n, ...) / tci_dmat_fetch move it), and a tci_lu made by tci_zrrlu is a complex
 * factorisation that only the tci_z* accessors accept.  Arithmetic that decides pivots is Julia Base's complex
 * arithmetic (include/tci_zarith.h: abs2, hypot, the robust division, multiply-then-subtract).                        */
/* rrlu(A::Matrix{ComplexF64}; maxrank, reltol, abstol, leftorthogonal), matrixlu.jl:141-225; outputs as tci_rrlu
 * (pivoterrors are the |.| of the pivots, real).                                                                      */
int tci_zrrlu(tci_ctx *ctx, const double *A_host, tci_dmat *A_dev, int64_t m, int64_t n, int64_t maxrank,
              double reltol, double abstol, int leftorthogonal, int64_t *rowperm, int64_t *colperm,
              int64_t *npivot, double *error, double *pivoterrors, tci_lu **factors /* nullable */);
/* lu.L (m x r) / lu.U (r x n), matrixlu.jl:374-392; left(luci) / right(luci), matrixluci.jl:40-84; B * A^-1 of
 * setsitetensor!, tensorci2.jl:391 -- for complex factors; matrices interleaved as above.                            */
int tci_zlu_fetch(tci_lu *lu, double *L /* nullable */, double *U /* nullable */);
int tci_zluci_left(tci_lu *lu, double *out_host /* nullable */, tci_dmat **out_dev /* nullable */);
int tci_zluci_right(tci_lu *lu, double *out_host /* nullable */, tci_dmat **out_dev /* nullable */);
int tci_zlu_rdiv(tci_lu *lu, tci_dmat *B, double *out_host /* nullable */, tci_dmat **out_dev /* nullable */);
/* Contraction(a, b) of two TensorTrain{ComplexF64,4} (contraction.jl:35-62) and TTCache(tt) of a
 * TensorTrain{ComplexF64,3} (cachedtensortrain.jl:9-30); cores interleaved.  tci_target_set_elementwise accepts
 * TCI_F_AFFINE with real a, b on them (the reference's tests use x -> 2x); tci_target_destroy releases them.          */
int tci_zmpo_pair_create(tci_ctx *ctx, int64_t nsites, const int64_t *dimsA4, const double *const *A,
                         const int64_t *dimsB4, const double *const *B, int64_t *target_id);
int tci_ztt_create(tci_ctx *ctx, int64_t nsites, const int64_t *dims3, const double *const *cores,
                   int64_t *target_id);
/* tci_pi_eval / tci_target_eval on a ComplexF64 target (contraction.jl:189-335, cachedtensortrain.jl:130-215):
 * out_host holds 2*nI*C*nJ doubles, out_dev is (2*nI*C) x nJ; *maxabs = max(abs.(out)) with abs = hypot.              */
int tci_zpi_eval(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                 int64_t nr, int64_t nJ, int64_t M, double *out_host, tci_dmat **out_dev, double *maxabs);
int tci_ztarget_eval(tci_ctx *ctx, int64_t target_id, const int64_t *idx, int64_t count, double *out /* 2*count */);
/* The `:full` branch of updatepivots! (tensorci2.jl:529-551) on a ComplexF64 target: Pi-eval -> rrLU, Pi never leaves
 * HBM.  Outputs as tci_bond_update.                                                                                   */
int tci_zbond_update(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                     int64_t nr, int64_t nJ, int64_t maxrank, double reltol, double abstol, int leftorthogonal,
                     int64_t *rowperm, int64_t *colperm, int64_t *npivot, double *error, double *pivoterrors,
                     double *maxabs, tci_lu **factors);
/* C = op(A) * op(B) on host Matrix{ComplexF64} arrays (plain transposes, as `_contract` permutes without conjugating,
 * contraction.jl:71-93): the complex FP64 tensor-core GEMM (four DMMA per complex tile) behind the chains above.      */
int tci_zgemm_host(tci_ctx *ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const double *A,
                   const double *B, double *C);
/* tci_contract_zipup_site / tci_contract_naive_site (contraction.jl:455-464, 338-349) on ComplexF64 cores; C_dev is the
 * (2*chi*s1*s3) x (Da'*Db') device matrix tci_zrrlu factorises next.                                                  */
int tci_zcontract_zipup_site(tci_ctx *ctx, const double *R, int64_t chi, int64_t Da, int64_t Db, const double *A,
                             int64_t s1, int64_t s2, int64_t Dan, const double *B, int64_t s3, int64_t Dbn,
                             double *C_host /* nullable */, tci_dmat **C_dev /* nullable */);
int tci_zcontract_naive_site(tci_ctx *ctx, const double *A, int64_t Da, int64_t s1, int64_t s2, int64_t Dan,
                             const double *B, int64_t Db, int64_t s3, int64_t Dbn, double *out_host);

#ifdef __cplusplus
}
#endif
#endif /* TCI_B200_H */
