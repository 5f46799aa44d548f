/*
 * vch_b200.h — C ABI of the B200 (sm_100a) forward/adjoint/PGD engine for sparse optimal control of the
 * viscous Cahn–Hilliard system.
 *
 * Every entry point replaces one Python function of the reference (paths relative to the reference's
 * src/ directory, cited per function).  The reference has no FFI of its own (it is pure Python); the
 * binding a maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; fp64 everywhere; arrays are dense C-order exactly as the reference's
 *     NumPy arrays: fields (Nx+1, Ny+1) with y contiguous, trajectories (levels, Nx+1, Ny+1), 1D (levels, N+1),
 *     ensembles carry a leading batch axis.
 *   - `mem` says where EVERY array argument of that call lives: VCH_MEM_HOST (the library stages the copies
 *     on its stream) or VCH_MEM_DEVICE (caller-owned device pointers, nothing is copied).  Small per-level
 *     vectors documented as "host" (time grids, coordinate vectors, scalars out) are always host memory.
 *   - every call is enqueued on the context's stream (vch_set_stream) and returns after the results
 *     it promises on the host are valid.  Return value: 0 = ok, otherwise a VCH_E_* code; vch_last_error()
 *     gives the message.  There is no CPU fallback: without a CUDA device every compute call fails
 *     with VCH_E_CUDA.
 *   - a context is not re-entrant: one context per host thread / stream.
 */
#ifndef VCH_B200_H
#define VCH_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VCH_MEM_HOST   0
#define VCH_MEM_DEVICE 1

#define VCH_OK            0
#define VCH_E_CUDA        1   /* CUDA runtime error / no device */
#define VCH_E_SHAPE       2   /* bad sizes or NULL where an array is required (reference: ValueError / AssertionError) */
#define VCH_E_NONFINITE   3   /* non-finite residual or mass defect (reference: RuntimeError, 1D Forward_solver.py:166-170) */
#define VCH_E_KRYLOV      4   /* linear solve stalled above tolerance (the reference's direct solve has no analogue) */
#define VCH_E_ARG         5
#define VCH_E_COMM        6   /* slab mode: a peer rank did not arrive within the wait limit */
#define VCH_IPC_HANDLE_BYTES 64

typedef struct vch2d_ctx vch2d_ctx;
typedef struct vch1d_ctx vch1d_ctx;

/* Physical/grid parameters.  2D: 2D/Vch_control_2D/config.py:103-113; delta_sep: Forward2_solver.py:510.
 * 1D: 1D/Vch_control_1D/config.py:93-102; delta_sep: Forward_solver.py:42. */
typedef struct {
    int    Nx, Ny;          /* intervals; nodes are (Nx+1) x (Ny+1) */
    double hx, hy;          /* spacings as the caller computed them (Lx/Nx or x[1]-x[0]) */
    double Lx, Ly;          /* only used by the (rare) uniform mass-correction fallback, Forward2_solver.py:576 */
    double tau, gamma, c1, c2, kappa;
    double delta_sep;
} vch2d_params;

typedef struct {
    int    N;               /* intervals; N+1 nodes */
    double h, Lx;
    double tau, gamma, c1, c2, kappa;
    double delta_sep;
} vch1d_params;

/* Work counters filled by the solvers (all optional: pass NULL). */
typedef struct {
    long long newton_residual_evals;   /* residual evaluations incl. line-search trials */
    long long newton_linear_solves;
    long long krylov_iterations;       /* BiCGStab iterations (2 operator + 2 preconditioner applies each) */
    long long krylov_max_iterations;   /* worst single solve */
    long long kernel_launches;         /* kernels this library launched during the call */
    long long krylov_stalls;           /* solves that stopped above tolerance */
    double    last_newton_residual;
    long long krylov_half_exits;       /* solves that ended after the first half of a BiCGStab iteration (counted as one iteration) */
    long long krylov_stalls_adjoint;   /* stalled solves of the ADJOINT sweep: vch2d_adjoint / vch2d_pgd_iteration return VCH_E_KRYLOV
                                          when > 0 (the forward Newton loop checks the true residual itself; the adjoint recurrence,
                                          backward2_solver.py:226-231, has no outer check, so an inaccurate p_n would silently
                                          corrupt every earlier level and the gradient) */
} vch_stats;

const char* vch_last_error(void);
int  vch_device_count(void);           /* 0 when no usable CUDA device */
int  vch_version(void);

/* ------------------------------------------------------------------ 2D context */
int  vch2d_create(const vch2d_params* p, int device, vch2d_ctx** out);
void vch2d_destroy(vch2d_ctx* c);

/* ---- slab mode: ONE 2D problem decomposed into row slabs over 2/4/8 GPUs of one node, one process per GPU
 * (BASELINE config 5).  The grid must be square with N = 2^k (32..4096).  Rank r owns rows
 * [r*N/R, (r+1)*N/R) of every (N+1, N+1) field (the last rank one more); every array argument of the vch2d_* calls
 * below is then the rank's slab (levels, rows, N+1), except x (cost / pgd_iteration), which stays the global abscissa
 * vector.  All ranks must make the same calls in the same order (the kernels exchange halo rows, transposes of the
 * spectral preconditioner and reduction partials through each other's memory over NVLink; there is no host-side
 * collective).  Setup: every rank creates its context, publishes the 64-byte IPC handle of its arena, collects all
 * ranks' handles (any host transport, e.g. torch.distributed.all_gather) and attaches.  vch2d_residual is not
 * available in slab mode.  A rank that waits more than 10 s for a peer fails with VCH_E_COMM instead of hanging. */
int  vch2d_slab_create(const vch2d_params* p_global, int device, int rank, int nranks, vch2d_ctx** out);
int  vch2d_slab_rows(vch2d_ctx* c, int* row0_out, int* nrows_out);
int  vch2d_slab_ipc_handle(vch2d_ctx* c, void* handle_out /* VCH_IPC_HANDLE_BYTES */);
int  vch2d_slab_attach(vch2d_ctx* c, const void* handles /* nranks * VCH_IPC_HANDLE_BYTES, rank order */);
int  vch2d_slab_selftest(vch2d_ctx* c, double* out5 /* sum, min, max of (rank+1); lower / upper ghost value */);
int  vch2d_set_stream(vch2d_ctx* c, void* cuda_stream);          /* cudaStream_t; NULL = legacy default stream */
int  vch2d_set_krylov(vch2d_ctx* c, double rel_tol, int max_iter);/* defaults 1e-11, 200 */
/* Forcing term of the time loop (vch2d_forward, vch2d_pgd_iteration): the FIRST linear solve of every Newton solve stops at
 * rel_tol (default 1e-6), the later ones at the vch2d_set_krylov tolerance.  The reference's direct solve
 * (Forward2_solver.py:367-372) has no analogue; Newton always iterates again after the first solve and that iteration
 * re-solves to full accuracy, so trajectories move by <= 2e-12 relative while the forward sweep needs 22 % fewer BiCGStab
 * iterations.  rel_tol = 0 switches it off (every solve to the full tolerance).  vch2d_newton / vch2d_jacobian_solve (the
 * test-level calls, whose residual histories are compared with the reference's) always solve to the full tolerance. */
int  vch2d_set_krylov_first(vch2d_ctx* c, double rel_tol);
/* Newton stop rule.  floor_aware = 1 (default): besides the reference's ||R||_2 < 1e-6 (Forward2_solver.py:353-365) the
 * iteration also stops when ||R||_2 reaches the fp64 resolution of the residual, eps*(1/hx^2+1/hy^2)*||mu||_2, or stalls within 50x of it — only
 * reachable on grids >= 1024^2, where the reference's absolute tolerance lies below that resolution and its loop would spin
 * to max_iter = 500 on rounding noise.  floor_aware = 0: the reference's rule verbatim. */
int  vch2d_set_newton(vch2d_ctx* c, int floor_aware);
/* Device memory the HOST-buffer path of vch2d_pgd_iteration may use for staging.  0 (default): stage the whole
 * trajectories when they fit, otherwise fall back to chunk rings sized to the free memory.  > 0: always walk the
 * trajectories in chunks through rings of at most `bytes` in total (trajectories larger than HBM; results agree with the
 * fully staged path to rounding of the cost sums). */
int  vch2d_set_stream_budget(vch2d_ctx* c, long long bytes);
long long vch2d_launch_count(vch2d_ctx* c);                        /* kernels launched since creation */
/* Per-kernel device timing (bench.py's roofline leg).  While enabled every launch is bracketed by CUDA events on the
 * launching stream; the report returns ';'-joined kernel names with total milliseconds and launch counts. */
int  vch2d_profile(vch2d_ctx* c, int enable);
int  vch2d_profile_report(vch2d_ctx* c, char* names, int names_cap, double* ms, long long* counts, int cap, int* n_out);

/* ------------------------------------------------------------------ 2D building blocks (test-level surface) */
/* apply_laplacian(L, v, Nx, Ny)                       2D/Vch_control_2D/Forward2_solver.py:140-152 (operator :105-137) */
int vch2d_apply_laplacian(vch2d_ctx* c, const double* v, double* out, int mem);
/* initialize_mu(phi, w, ...)                          Forward2_solver.py:155-167 */
int vch2d_initialize_mu(vch2d_ctx* c, const double* phi, const double* w, double* mu_out, int mem);
/* solve_w(w_old, dt, gamma, u_n, u_np1)               Forward2_solver.py:170-181  (count elements; 1D and 2D) */
int vch_solve_w(void* cuda_stream, long long count, const double* w_old, double dt, double gamma,
                const double* u_n, const double* u_np1, double* w_new_out, int mem);
/* solve_phi_residual + solve_mu_residual              Forward2_solver.py:184-221 */
int vch2d_residual(vch2d_ctx* c, const double* phi_new, const double* phi_old, const double* mu_new,
                   const double* mu_old, const double* w_new, const double* w_old, double dt,
                   double* Rphi_out, double* Rmu_out, int mem);
/* spsolve(assemble_jacobian(phi), -R)                 Forward2_solver.py:224-253, :367-372
 * matrix-free: Schur reduction + left-preconditioned BiCGStab with the DCT-I fast solve */
int vch2d_jacobian_solve(vch2d_ctx* c, const double* phi, double dt, const double* Rphi, const double* Rmu,
                         double* dphi_out, double* dmu_out, int* krylov_iters_out, int mem);
/* newton_raphson(...)                                 Forward2_solver.py:323-427
 * res_hist (host, capacity hist_cap) receives ||R||_2 per Newton iteration, *n_hist their number */
int vch2d_newton(vch2d_ctx* c, const double* phi_old, const double* mu_old, const double* w_old,
                 const double* w_new, double dt, double* phi_new_out, double* mu_new_out,
                 double* res_hist, int hist_cap, int* n_hist, vch_stats* stats, int mem);

/* ------------------------------------------------------------------ 2D hot path */
/* run_main_simulation time loop                       Forward2_solver.py:542-585
 *   phi0: initial field (init_phi_random is host RNG, :444-486, generated by the caller)
 *   dt_steps (host, n_steps): the caller's min(dt, T-t) sequence (:543)
 *   u: control (u_rows, Nx+1, Ny+1) or NULL; step s reads rows s and s+1 while s < u_rows-1, zeros after (:545-548)
 *   phi_hist_out: (n_steps+1, ...) incl. level 0;  mu_hist_out / w_hist_out: (n_steps, ...) or NULL */
int vch2d_forward(vch2d_ctx* c, const double* phi0, const double* u, int u_rows, int n_steps,
                  const double* dt_steps, double* phi_hist_out, double* mu_hist_out, double* w_hist_out,
                  vch_stats* stats, int mem);
/* run_backward                                        2D/Vch_control_2D/backward2_solver.py:75-246
 *   t_hist (host, levels); phiQ (levels, ...) or NULL; phiT (field) or NULL; p_out/q_out may be NULL */
int vch2d_adjoint(vch2d_ctx* c, const double* phi_hist, int levels, const double* t_hist, double b1, double b2,
                  const double* phiQ, const double* phiT, double* p_out, double* q_out, double* r_out,
                  vch_stats* stats, int mem);
/* calculate_cost                                      2D/Vch_control_2D/cost2_and_function.py:80-108
 *   x (host, Nx+1), y (host, Ny+1), t_hist (host, levels): np.trapz abscissae
 *   J_out (host, 9) = J, J1..J4, then the raw trapezoid integrals  int|phi-phi_Q|^2, int|phi(T)-phi_T|^2, int u^2, int|u|
 *   (the driver's tracking / terminal error norms, GD2_configured.py:336-363, are square roots of the first two) */
int vch2d_cost(vch2d_ctx* c, const double* phi_hist, const double* u, const double* phiQ, const double* phiT,
               int levels, const double* x, const double* y, const double* t_hist,
               double b1, double b2, double b3, double kappa_sp, double* J_out, int mem);
/* calculate_gradient + proximal_step (+ driver norms) cost2_and_function.py:150, :191-200; GD2_configured.py:375
 * 1D: cost_and_function.py:99, :111; GD_1D.py:56-71, :466.   grad_out may be NULL.
 *   red_out (host, 4) = { ||u_new-u||^2, ||u||^2, #(u_new != 0), #(u_new at a bound) } */
int vch_grad_prox(void* cuda_stream, long long count, const double* u, const double* r, double b3, double alpha,
                  double kappa_sp, double u_min, double u_max, double* grad_out, double* u_new_out,
                  double* red_out, int mem);
/* verify_sparsity_condition counts                    second_order_conditions_2d.py:238-297; GD_1D.py:115-147
 *   counts_out (host, 3) = { #|u|<tol, #|r|<=kappa, #agree } */
int vch_kkt_counts(void* cuda_stream, long long count, const double* u, const double* r, double kappa_sp,
                   double tol, long long* counts_out, int mem);
/* free_energy(phi, kappa, c1, c2, hx, hy, w, eps)      Forward2_solver.py:256-319 (monitoring)
 *   phi (n0, n1) row-major = the reference's (Ny+1, Nx+1): h1 = hx is the spacing of the contiguous axis, h0 = hy of the other;
 *   w may be NULL; one fused reduction kernel; E_out (host, 1) */
int vch_free_energy(void* cuda_stream, int n0, int n1, const double* phi, const double* w, double kappa, double c1, double c2,
                    double h1, double h0, double eps, double* E_out, int mem);
/* One optimistic PGD iteration                        GD2_configured.py:299-313
 *   adjoint(phi_hist) -> r;  u_new = prox(u - alpha (r + b3 u));  forward(u_new) -> phi_hist_out;  J(u_new).
 *   levels = n_steps+1 = rows of u, phi_hist, phiQ.  phi0 = phi_hist level 0.  r_out may be NULL (device scratch is used).
 *   J_out (host, 9), red_out (host, 4) as above. */
int vch2d_pgd_iteration(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps,
                        const double* x, const double* y,
                        const double* u, const double* phi_hist, const double* phiQ, const double* phiT,
                        double b1, double b2, double b3, double kappa_sp, double u_min, double u_max, double alpha,
                        double* u_new_out, double* phi_hist_out, double* r_out,
                        double* J_out, double* red_out, vch_stats* stats, int mem);

/* ---- trajectories that do not fit in HBM: checkpoint + recompute (BASELINE north_star (2); the reference keeps the whole
 * trajectory in RAM, Forward2_solver.py:535-537, :583-585).  The state trajectory exists only as checkpoints (phi, mu, w) at the
 * levels 0, S, 2S, ..., M (ncp = ceil(M / S) + 1 entries each, S = ckpt_stride, M = levels - 1):
 *   ck_phi[j] = phi at level min(jS, M);  ck_mu[j] / ck_w[j] = mu / w after that many steps (entry 0: initialize_mu(phi_0, 0) and 0).
 * vch2d_forward_ckpt      run_main_simulation that keeps only the checkpoints.
 * vch2d_pgd_iteration_ckpt one optimistic PGD iteration (GD2_configured.py:299-313): the adjoint sweep walks the segments from the last
 *   to the first, recomputing each from its checkpoint (same kernels, same order, deterministic reductions: the recomputed levels
 *   are bit-identical to the ones the checkpoints came from) and applying gradient + prox to its levels at once; the forward solve
 *   under the new control writes the new checkpoints and accumulates the cost per segment.  u, phiQ, u_new_out stay full
 *   trajectories (levels, ...); r_out may be NULL (then r exists only for S + 1 levels at a time).  Results equal those of
 *   vch2d_pgd_iteration on the fully stored trajectory bit for bit (u_new, r, checkpoints) resp. to the rounding of the segmented
 *   cost sums (J).  Cost: one additional forward sweep per iteration; memory: 6 (M/S + 1) + 2 (S + 1) levels instead of 3 (M + 1). */
int vch2d_forward_ckpt(vch2d_ctx* c, const double* phi0, const double* u, int u_rows, int n_steps, const double* dt_steps,
                       int ckpt_stride, double* ck_phi_out, double* ck_mu_out, double* ck_w_out, vch_stats* stats, int mem);
int vch2d_pgd_iteration_ckpt(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps,
                             const double* x, const double* y, const double* u,
                             const double* ck_phi, const double* ck_mu, const double* ck_w, int ckpt_stride,
                             const double* phiQ, const double* phiT,
                             double b1, double b2, double b3, double kappa_sp, double u_min, double u_max, double alpha,
                             double* u_new_out, double* ck_phi_out, double* ck_mu_out, double* ck_w_out, double* r_out,
                             double* J_out, double* red_out, vch_stats* stats, int mem);

/* ------------------------------------------------------------------ 1D (batched ensembles; batch = 1 is the reference call) */
int  vch1d_create(const vch1d_params* p, int device, vch1d_ctx** out);
void vch1d_destroy(vch1d_ctx* c);
int  vch1d_set_stream(vch1d_ctx* c, void* cuda_stream);
long long vch1d_launch_count(vch1d_ctx* c);
/* solve_phi_residual + solve_mu_residual              1D/Vch_control_1D/Forward_solver.py:93-109  (batch rows) */
int vch1d_residual(vch1d_ctx* c, int batch, const double* phi_new, const double* phi_old, const double* mu_new,
                   const double* mu_old, const double* w_new, const double* w_old, double dt,
                   double* Rphi_out, double* Rmu_out, int mem);
/* initialize_mu(phi, w, c1, c2, L, kappa)              Forward_solver.py:82-86 */
int vch1d_initialize_mu(vch1d_ctx* c, int batch, const double* phi, const double* w, double* mu_out, int mem);
/* newton_raphson                                      Forward_solver.py:139-235.  res_hist (host, batch*hist_cap), n_hist (host, batch) */
int vch1d_newton(vch1d_ctx* c, int batch, const double* phi_old, const double* mu_old, const double* w_old,
                 const double* w_new, double dt, double* phi_new_out, double* mu_new_out,
                 double* res_hist, int hist_cap, int* n_hist, int* status_out, int mem);
/* run_main_simulation time loop                       Forward_solver.py:342-374
 *   phi0 (batch, N+1); u (batch, u_rows, N+1) or NULL — step s reads rows s and s+1, the last row repeats (:347-353)
 *   phi_hist_out (batch, n_steps+2, N+1): level 0 is stored twice (:329-336); mu/w hist (batch, n_steps, N+1) or NULL
 *   status_out (host, batch) or NULL: per-problem VCH_OK / VCH_E_NONFINITE */
int vch1d_forward(vch1d_ctx* c, int batch, const double* phi0, const double* u, int u_rows, int n_steps,
                  const double* dt_steps, double* phi_hist_out, double* mu_hist_out, double* w_hist_out,
                  int* status_out, vch_stats* stats, int mem);
/* run_backward                                        1D/Vch_control_1D/backward_solver.py:48-125
 *   physics come from the context (the reference hard-wires ForwardSolverConfig() defaults, :29-33);
 *   b1, b2 (host, batch); t_hist (host, levels) shared; levels with dt<=0 are skipped leaving zeros (:110) */
int vch1d_adjoint(vch1d_ctx* c, int batch, const double* phi_hist, int levels, const double* t_hist,
                  const double* b1, const double* b2, const double* phiQ, const double* phiT,
                  double* p_out, double* q_out, double* r_out, int mem);
/* calculate_cost                                      1D/Vch_control_1D/cost_and_function.py:55-75
 *   weights (host, batch*4) = b1,b2,b3,kappa_sp per problem; J_out (host, batch*5) */
int vch1d_cost(vch1d_ctx* c, int batch, const double* phi_hist, const double* u, const double* phiQ,
               const double* phiT, int levels, const double* x, const double* t_hist,
               const double* weights, double* J_out, int mem);
/* gradient + prox with per-problem parameters (ensembles).  par (host, batch*6) = b3, alpha, kappa_sp, u_min, u_max, unused
 *   per_problem = levels*(N+1) elements;  red_out (host, batch*4) */
int vch1d_grad_prox(vch1d_ctx* c, int batch, long long per_problem, const double* u, const double* r,
                    const double* par, double* u_new_out, double* red_out, int mem);

#ifdef __cplusplus
}
#endif
#endif /* VCH_B200_H */
