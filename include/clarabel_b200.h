/* clarabel_b200.h — C-ABI of the B200-native KKT linear-system backend for Clarabel.jl.
 *
 * Drop-in boundary.  Two reference interfaces are covered:
 *
 *  (1) INNER = the reference's LDL plugin API `AbstractDirectLDLSolver`
 *      (/root/reference/src/kktsolvers/direct-ldl/directldl_defaults.jl:1-72; the QDLDL
 *      implementation it replaces is directldl_qdldl.jl:1-96).  One C function per method.
 *  (2) OUTER = `AbstractKKTSolver` (/root/reference/src/kktsolvers/kktsolver_defaults.jl:1-47,
 *      concrete reference implementation kktsolver_directldl.jl), fused so that cone state is
 *      uploaded once per iteration and the cone->K update, static regularisation,
 *      factorisation, solves and iterative refinement all run on the device.
 *
 * Conventions: every function returns int32 status: 0 = ok, >0 = numerical failure (the Julia
 * shim returns `false`, never throws — kktsolver_directldl.jl:281,356-370), <0 = usage/CUDA error
 * (the shim calls `error(...)`).  Host pointers in/out, synchronous on return (the reference
 * caller is single-threaded and synchronous).  Indices are 0- or 1-based according to
 * cb200_settings.index_base.  Float64 only (T == Float64, as for every external reference
 * engine: ext/directldl_hsl.jl:18).
 */
#ifndef CLARABEL_B200_H
#define CLARABEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb200_handle cb200_handle;

/* Mirrors the KKT-relevant fields of `Settings{Float64}` (src/settings.jl:110-132). */
typedef struct cb200_settings {
    int32_t index_base;                 /* 0 (C / Python harness) or 1 (Julia) */
    int32_t device;                     /* CUDA device ordinal */
    int32_t static_regularization_enable;         /* settings.jl:117 */
    double  static_regularization_constant;       /* :118  (1e-8) */
    double  static_regularization_proportional;   /* :119  (eps^2) */
    int32_t dynamic_regularization_enable;        /* :122 */
    double  dynamic_regularization_eps;           /* :123  (1e-13) */
    double  dynamic_regularization_delta;         /* :124  (2e-7) */
    int32_t iterative_refinement_enable;          /* :127 */
    double  iterative_refinement_reltol;          /* :128  (1e-13) */
    double  iterative_refinement_abstol;          /* :129  (1e-12) */
    int32_t iterative_refinement_max_iter;        /* :131  (10) */
    double  iterative_refinement_stop_ratio;      /* :132  (5) */
    int32_t ordering;                   /* 0 AMD, 1 auto (default): nested dissection unless > 3x AMD cost;
                                         * AMD when K holds a dense PSD cone block (detected from pattern + Dsigns), 2 natural */
    double  amd_dense_scale;            /* dense-row threshold multiplier (reference QDLDL path: 1.5, directldl_qdldl.jl:24; default here 0.3) */
    int32_t nd_leaf_size;
    int32_t use_cuda_graph;             /* bit mask: 1 replay the solve sweeps, 2 the factorisation, through CUDA graphs */
    int32_t reserved[8];
} cb200_settings;

void cb200_default_settings(cb200_settings* s);

/* ---------------------------------------------------------------- host-only symbolic analysis
 * (no CUDA call; used by tests and by the CPU-side tooling).  Replaces the symbolic half of
 * QDLDL.qdldl(KKT; logical=true) — directldl_qdldl.jl:18-25. */
typedef struct cb200_symbolic cb200_symbolic;
int32_t cb200_symbolic_create(int64_t N, const int64_t* colptr, const int64_t* rowval,
                              const cb200_settings* st, const int64_t* user_perm,
                              cb200_symbolic** out);
void    cb200_symbolic_destroy(cb200_symbolic* s);
/* what: 0 N, 1 nsuper, 2 nnzL, 3 nlevels, 4 max_front, 5 max_width, 6 upd_total, 7 panel_total,
 *       8 rows_total, 9 nnzK */
int64_t cb200_symbolic_stat(const cb200_symbolic* s, int32_t what);
double  cb200_symbolic_flops(const cb200_symbolic* s);
/* which: 0 perm[N], 1 sn_first[nsuper+1], 2 rows_ptr[nsuper+1], 3 rows[rows_total],
 *        4 rel[rows_total], 5 sn_parent[nsuper], 6 panel_off[nsuper+1], 7 upd_off[nsuper+1],
 *        8 a_map[nnzK], 9 sn_level[nsuper], 10 child_ptr[nsuper+1], 11 child_list[nsuper-#roots]
 * All arrays are returned widened to int64 into caller storage of the stated length. */
int32_t cb200_symbolic_get(const cb200_symbolic* s, int32_t which, int64_t* out, int64_t len);

/* Multi-GPU partition of the assembly tree (host-only): owner[s] in [0,nranks) or -1 for the
 * replicated top set, is_top[s] in {0,1}, rank_load[nranks] = factor-flop weight per rank. */
int32_t cb200_symbolic_partition(const cb200_symbolic* s, int32_t nranks, int64_t* owner,
                                 int64_t* is_top, double* rank_load);

/* Optional ordering hint, consumed by the NEXT cb200_symbolic_create / cb200_create on this thread:
 * block_id[i] >= 0 marks row/column i of K as part of a dense cone block (PSD triangle, dense SOC);
 * nested dissection then keeps each block in one piece (a separator cutting such a clique was
 * observed to destabilise the factorisation late in the IP iteration).  NULL clears the hint. */
int32_t cb200_hint_blocks(const int64_t* block_id, int64_t N);

/* Fill-reducing orderings (perm[k] = original index of the k-th pivot). */
int32_t cb200_order_amd(int64_t n, const int64_t* colptr, const int64_t* rowval,
                        double dense_scale, int64_t* perm);
int32_t cb200_order_nd(int64_t n, const int64_t* colptr, const int64_t* rowval,
                       double dense_scale, int64_t leaf_size, int64_t* perm);

/* ---------------------------------------------------------------- INNER boundary
 * Constructor `MyT{T}(KKT::SparseMatrixCSC, Dsigns::Vector{Int}, settings)` called at
 * kktsolver_directldl.jl:86: symbolic analysis + device upload.  `colptr/rowval/nzval` is the
 * upper-triangular (:triu) KKT matrix; Dsigns in {+1,-1}. */
int32_t cb200_create(int64_t N, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                     const int64_t* Dsigns, const cb200_settings* st, cb200_handle** out);
void    cb200_destroy(cb200_handle* h);                       /* Julia finalizer (MOI_wrapper.jl:133) */
/* update_values!(s, index, values)   — directldl_qdldl.jl:46-56 */
int32_t cb200_update_values(cb200_handle* h, const int64_t* index, const double* values, int64_t len);
/* scale_values!(s, index, scale)     — directldl_qdldl.jl:60-68 */
int32_t cb200_scale_values(cb200_handle* h, const int64_t* index, int64_t len, double scale);
/* refactor!(s, KKT)::Bool            — directldl_qdldl.jl:72-81 ; >0 = non-finite pivot */
int32_t cb200_refactor(cb200_handle* h);
/* solve!(s, KKT, x, b)               — directldl_qdldl.jl:85-96 */
int32_t cb200_solve(cb200_handle* h, double* x, const double* b);
/* linear_solver_info(s)              — directldl_qdldl.jl:35-42 */
int32_t cb200_info(const cb200_handle* h, int64_t* nnzA, int64_t* nnzL, int32_t* ngpus);

/* ---------------------------------------------------------------- OUTER boundary (fused)
 * cb200_set_maps: once after create; uploads LDLDataMap (directldl_datamaps.jl:170-214) and the
 * cone table (CompositeCone.rng_cones / rng_blocks, compositecone_type.jl:96-141).
 *   cone_type[i]: 0 Zero, 1 Nonnegative, 2 SecondOrder, 3 PSDTriangle, 4 Exponential, 5 Power,
 *                 6 GenPower ; cone_dim[i]: dim (PSD: side n)
 *   map_soc_u / map_soc_v: concatenated over sparse SOCs (dim > 4) in cone order; map_soc_D: 2 each.
 *   Nonsymmetric cones (types 4-6): their Hs block (3x3 for exp/pow, coneops_expcone.jl:92-100;
 *   diagonal + expansion columns for genpow, directldl_datamaps.jl:146-166) is the output of the
 *   caller's update_scaling!; the caller sends those few values with cb200_update_values right
 *   before cb200_update_cones, which leaves the entries untouched. */
int32_t cb200_set_maps(cb200_handle* h, int64_t n, int64_t m, int64_t p,
                       const int64_t* map_P, int64_t nnzP, const int64_t* map_A, int64_t nnzA,
                       const int64_t* map_Hs, int64_t nHs, const int64_t* map_diag_full,
                       int64_t ncones, const int32_t* cone_type, const int64_t* cone_dim,
                       const int64_t* map_soc_u, const int64_t* map_soc_v, const int64_t* map_soc_D);
/* kktsolver_update!(ks, cones)::Bool — kktsolver_directldl.jl:197-294 (get_Hs!, sign flip,
 * sparse-cone scatter, static regularisation, refactor).  Cone scaling state as produced by
 * update_scaling! on the host: w over all m rows (NN: w, SOC: w), per-SOC eta and d, SOC u/v
 * over all SOC rows, PSD R matrices (column-major n*n each, cone order). */
int32_t cb200_update_cones(cb200_handle* h, const double* w, const double* soc_eta,
                           const double* soc_d, const double* soc_u, const double* soc_v,
                           const double* psd_R);
/* kktsolver_setrhs!(ks, rhsx, rhsz) — kktsolver_directldl.jl:313-327: b = [rhsx; rhsz; 0_p].  The
 * right-hand side is copied to the device here (the caller may reuse its buffers on return);
 * a following cb200_solve_ir with rhsx == rhsz == NULL solves for it. */
int32_t cb200_setrhs(cb200_handle* h, const double* rhsx, const double* rhsz);
/* kktsolver_setrhs! + kktsolver_solve! (+ getlhs) — kktsolver_directldl.jl:313-371, with
 * _iterative_refinement :389-449 on the device.  lhsx / lhsz may be NULL (Julia `nothing`);
 * rhsx == rhsz == NULL: use the right-hand side set by cb200_setrhs. */
int32_t cb200_solve_ir(cb200_handle* h, const double* rhsx, const double* rhsz,
                       double* lhsx, double* lhsz, int32_t* ir_rounds);
/* kktsolver_update_P!/A! — kktsolver_directldl.jl:374-386 */
int32_t cb200_update_P(cb200_handle* h, const double* values, int64_t len);
int32_t cb200_update_A(cb200_handle* h, const double* values, int64_t len);

/* ---------------------------------------------------------------- introspection (tests / bench)
 * what: 0 device KKT nzval (unregularised, len nnzK), 1 D (permuted, len N), 2 panel storage,
 *       3 perm (as doubles), 4 last static regulariser (len 1), 5 regularised-pivot count (len 1),
 *       6 full solution [x; z; expansion variables] of the last solve (len N),
 *       7 original indices of the first 64 dynamically regularised pivots of the last factorisation (len 64, -1 = none) */
int32_t cb200_download(cb200_handle* h, int32_t what, double* out, int64_t len);
/* timers (ms, accumulated CUDA-event times on the handle's stream): 0 cone update + scatter,
 * 1 factor, 2 triangular solves, 3 spmv/residual ; counters: 4 #factor, 5 #solves, 6 #kernel launches ;
 * with cb200_set_detail(h,1) also per kernel class inside the factorisation: 7 k_schur_large,
 * 8 pivot-block phase (k_diag64 + k_rows64), 9 small fronts (k_factor_small), 10 assembly */
int32_t cb200_get_timers(cb200_handle* h, double* out, int32_t len);
int32_t cb200_reset_timers(cb200_handle* h);
/* event timing inside the phases (disables CUDA-graph replay while on): level 0 off, 1 the four
 * factorisation groups of cb200_get_timers, 2 additionally every kernel class of the factorisation
 * and of the solve sweeps (cb200_get_fine_timers) */
int32_t cb200_set_detail(cb200_handle* h, int32_t level);
/* detail level 2: accumulated ms per kernel class; writes min(len, count) values, returns count;
 * cb200_fine_timer_name(i) names class i.  Reset by cb200_reset_timers. */
int32_t cb200_get_fine_timers(cb200_handle* h, double* out_ms, int32_t len);
const char* cb200_fine_timer_name(int32_t i);
/* symbolic statistics: 0 factor flops (sum of squared column lengths), 1 flops of the large-front
 * Schur GEMMs, 2 flops of the large-front panel phase, 3 nnzL, 4 levels, 5 supernodes, 6 large
 * fronts, 7 panel bytes of the multi-CTA solve class, 8 update-storage bytes, 9 panel-storage bytes */
int32_t cb200_get_stats(const cb200_handle* h, double* out, int32_t len);
/* ---------------------------------------------------------------- multi-GPU (one process per GPU)
 * Every rank creates the same handle on its own device (replicated symbolic analysis), then
 * calls cb200_dist_init with a NCCL unique id obtained on rank 0 (cb200_nccl_unique_id) and
 * distributed by the caller (torch.distributed / MPI).  Independent elimination-tree subtrees
 * are factored and solved on their owner rank; the replicated top (root separator) fronts are
 * assembled with an NCCL all-reduce over NVLink.  All ranks must then issue the same sequence
 * of refactor / update_cones / solve calls with the same inputs; results are replicated. */
int32_t cb200_nccl_unique_id(char out[128]);
int32_t cb200_dist_init(cb200_handle* h, int32_t rank, int32_t nranks, const char uid[128]);

/* The CUDA stream (cudaStream_t) all of this handle's work is enqueued on, so a caller can
 * bracket calls with its own events. */
void*   cb200_get_stream(cb200_handle* h);
/* resident != 0: the inputs live in HBM.  Non-NULL pointer arguments of cb200_update_cones and
 * cb200_setrhs are then DEVICE pointers (copied device-to-device on the handle's stream), NULL
 * keeps what the previous call left on the device; cb200_solve_ir ignores its host pointers and
 * leaves the solution on the device (cb200_download selector 6).  This is the mode bench.py's
 * device-resident `value` is measured in; the drop-in path (resident == 0) takes host pointers. */
int32_t cb200_set_resident(cb200_handle* h, int32_t resident);
const char* cb200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CLARABEL_B200_H */
