/*
 * fgd_b200.h -- C ABI of the B200-native batched functional-gradient-descent
 * (FGD) trajectory optimiser.  libfgd_b200.so exports exactly these symbols.
 *
 * The reference (simongroeger/irm_motion_planning) has no FFI; the seam this
 * library replaces is the jitted operator
 *     jit_optimize(self, alpha, obstacles, start_config, goal_config) -> alpha
 *         optimizer_BLS.py:126-213, optimizer_GD.py:68-97 and :172-232
 * together with the objective it closes over
 *     Trajectory.compute_trajectory_cost / _g     trajectory.py:271-297
 *     Trajectory.constraintsFulfilled             trajectory.py:129-137
 * generalised from ONE trajectory to a batch of B independent trajectories.
 * Traced inputs of the reference operator (alpha, obstacles, start, goal) are
 * runtime buffers here; everything the reference closes over (K, dK, J, robot
 * and hyper-parameters) is fixed at fgd_create() time.
 *
 * Conventions
 *  - plain C types only; all pointers marked d_ are CUDA device pointers owned
 *    by the caller (PyTorch in the Python host), h_ are host pointers.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - every function returns 0 on success, an FgdStatus code otherwise; nothing
 *    calls exit() (the reference prints "FATAL" and exit(-1), the Python host
 *    re-creates that behaviour on top of these codes).
 *  - one host thread per handle; calls on one handle are not re-entrant.  Launches
 *    of one handle may be enqueued on different streams: every launch owns its
 *    work-queue counter (a ring of FGD_QUEUE_RING counters), so they may overlap on
 *    the device; at most FGD_QUEUE_RING optimise launches of one handle may be in
 *    flight at once.  The host-buffer entry points share one staging area and are
 *    synchronous.
 *  - all arithmetic is FP32, the reference's precision (JAX default, x64 off).
 */
#ifndef FGD_B200_H
#define FGD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGD_ABI_VERSION 3
#define FGD_MAX_T 256            /* time samples / RKHS support points            */
#define FGD_MAX_OUTER 16         /* length of the gd_lr table                     */
#define FGD_FSTATE 8             /* floats of resumable state per trajectory      */
#define FGD_ISTATE 8             /* int32s of resumable state per trajectory      */
#define FGD_QUEUE_RING 64        /* optimise launches of one handle that may be in flight at once */
#define FGD_OBS_RING 16          /* published obstacle sets kept on the device (power of two)      */
#define FGD_SWITCH_LOG 64        /* int32 pairs per trajectory in the optional switch log of fgd_optimize_live */

typedef enum FgdStatus {
    FGD_OK = 0,
    FGD_ERR_INVALID_ARGUMENT = 1,
    FGD_ERR_UNSUPPORTED_T = 2,        /* T < 2 or T > FGD_MAX_T                    */
    FGD_ERR_KERNEL_NOT_SYMMETRIC = 3, /* km != km^T or dkm != -dkm^T bit-wise      */
    FGD_ERR_TOO_MANY_OBSTACLES = 4,   /* count > obstacle_capacity, or a capacity whose staging buffer
                                         does not fit in shared memory (~20 000 obstacles)        */
    FGD_ERR_CUDA = 5,                 /* see fgd_last_cuda_error()                 */
    FGD_ERR_NO_DEVICE = 6,
    FGD_ERR_JOINTS = 7                /* n_joints != 3 (robot.py:31, trajectory.py:42) */
} FgdStatus;

/* fstate[b][FGD_FSTATE] */
enum { FGD_F_LAM_SG = 0, FGD_F_LAM_JL = 1, FGD_F_LR = 2, FGD_F_LOSS = 3, FGD_F_TOC = 4, FGD_F_LAST_NEW_LOSS = 5 };
/* istate[b][FGD_ISTATE] */
enum { FGD_I_STATUS = 0, FGD_I_OUTER = 1, FGD_I_INNER = 2, FGD_I_INNER_TOTAL = 3, FGD_I_CAND_EVALS = 4,
       FGD_I_ACCEPTS = 5, FGD_I_FULFILLED = 6, FGD_I_HASH = 7 };
/* istate[b][FGD_I_STATUS]; a zero-filled state means "fresh, start from lambda_*_constraint" */
enum { FGD_ST_FRESH = 0, FGD_ST_ACTIVE = 1, FGD_ST_DONE = 2 };

/* Static configuration = everything the reference's optimizer object closes
 * over (main.py:17-98 flag table). */
typedef struct FgdConfig {
    int32_t abi_version;                 /* FGD_ABI_VERSION                                  */
    int32_t n_timesteps;                 /* --n-timesteps            trajectory.py:34        */
    int32_t n_joints;                    /* --n-joints, must be 3    robot.py:18             */
    int32_t obstacle_capacity;           /* max obstacles the device buffer can hold         */
    int32_t strict_math;                 /* 1: IEEE reciprocal, division and square root - every result bit-identical to the CPU oracle
                                            (1.11-1.42 x the fast-math time); 0: rcp.approx, Markstein division */
    int32_t max_inner_iteration;         /* optimizer_BLS.py:27                              */
    int32_t max_outer_iteration;         /* optimizer_BLS.py:28                              */
    int32_t max_bls_iteration;           /* optimizer_BLS.py:39                              */
    int32_t constraint_violating_dependant_loss; /* trajectory.py:28                         */
    int32_t n_gd_lr;                     /* valid entries of gd_lr   optimizer_GD.py:34-38   */
    int32_t whole_arm_cost;              /* 0: the reference's end-effector obstacle cost (trajectory.py:113-126);
                                            1: summed over all joint positions fk_1..fk_3 (robot.py:39-72), the
                                            extension named in DevBlog-Theme/blog-post.html:505-513              */
    float lambda_sg_constraint, lambda_jl_constraint, lambda_constraint_increase;   /* optimizer_BLS.py:32-34 */
    float lambda_max_cost, lambda_reg;                                              /* optimizer_BLS.py:36-37 */
    float loop_loss_reduction;                                                      /* optimizer_BLS.py:30    */
    float eps_position, eps_velocity;                                               /* robot.py:25-26         */
    float bls_lr_start, bls_alpha, bls_beta_plus, bls_beta_minus;                   /* optimizer_BLS.py:40-43 */
    float joint_safety_limit;                                                       /* trajectory.py:29       */
    float max_joint_position, min_joint_position, max_joint_velocity;               /* robot.py:14-16         */
    float link_length[3];                                                           /* robot.py:19            */
    float jac[9];                        /* row-major 3x3            trajectory.py:42        */
    float gd_lr[FGD_MAX_OUTER];          /* --gd-lr                  optimizer_GD.py:38      */
    const float *h_km;                   /* host, row-major T x T    trajectory.py:40        */
    const float *h_dkm;                  /* host, row-major T x T    trajectory.py:41        */
} FgdConfig;

typedef struct FgdHandle FgdHandle;

/* Replaces the optimizer constructors' setup (optimizer_BLS.py:23-48,
 * optimizer_GD.py:15-45): copies K, dK (re-laid out), J and all scalars to the
 * current CUDA device.  The caller keeps ownership of the host arrays. */
int fgd_create(const FgdConfig *cfg, FgdHandle **out);
int fgd_destroy(FgdHandle *h);

const char *fgd_status_string(int status);
/* cudaError_t of the last failing CUDA call on this handle (0 if none). */
int fgd_last_cuda_error(const FgdHandle *h);

/* Replaces passing `self.env.obstacles` as a traced argument
 * (optimizer_BLS.py:60,79,82,90; environment.py:17-29): PUBLISHES `count` (x,y) pairs as the next obstacle
 * generation: the set is copied (cudaMemcpyAsync on `stream`, from an internal page-locked staging copy of a host
 * source, so the caller's buffer may be reused at once) into the next slot of a ring of FGD_OBS_RING device-resident
 * sets, then the slot header and the "latest generation" word are updated, in stream order.  Every launch enqueued
 * afterwards uses it; a RUNNING fgd_optimize_live kernel picks it up at its next poll.  No recompilation, no handle
 * re-creation.  xy_on_device: 0 = host pointer, 1 = device pointer (copied device-to-device; must stay valid until
 * the copy has run).  A slot is overwritten FGD_OBS_RING generations later; launches that captured it are waited for
 * (stream-ordered), a running live kernel validates what it copied against the slot header instead. */
int fgd_set_obstacles_async(FgdHandle *h, const float *xy, int32_t count, int32_t xy_on_device, void *stream);
int fgd_obstacle_count(const FgdHandle *h);
int fgd_obstacle_generation(const FgdHandle *h);   /* number of sets published so far (0 = the initial empty set) */

/* Unit-parity hook = compute_trajectory_cost + compute_trajectory_cost_g +
 * constraintsFulfilled for B trajectories at given penalty weights
 * (trajectory.py:271-297,129-137).  Any output pointer may be NULL.
 *   lambda_max_cost: the reference's static argument (main.py:141-142 calls it
 *                    with 0 and 1 for the final report); < 0 = the configured value
 *   d_alpha [B][T][3], d_start/d_goal [B][3]
 *   d_loss [B], d_toc [B] (obstacle term only), d_grad [B][T][3],
 *   d_q / d_v [B][T][3] (K alpha J, dK alpha J), d_fulfilled [B] (0/1) */
int fgd_eval_cost_grad(FgdHandle *h, int32_t B, const float *d_alpha, const float *d_start, const float *d_goal,
                       float lambda_sg, float lambda_jl, float lambda_max_cost, float *d_loss, float *d_toc,
                       float *d_grad, float *d_q, float *d_v, int32_t *d_fulfilled, void *stream);

/* Replace BacktrackingLineSearchOptimizer.jit_optimize (optimizer_BLS.py:126-213)
 * and GradientDescentOptimizer.jit_optimize / jit_dual_optimize
 * (optimizer_GD.py:68-97 when max_outer_iteration == 1, :172-232 otherwise)
 * for B trajectories.  d_alpha is updated in place.  d_fstate/d_istate carry the
 * loop state between calls (zero-fill for a fresh run): each call advances every
 * unfinished trajectory by at most `max_launch_iters` inner iterations
 * (< 0: run to completion), re-evaluating the loss at the current obstacle set
 * first -- the plain loop's "read self.env.obstacles afresh" semantics
 * (optimizer_BLS.py:79,82,90). */
int fgd_optimize_bls(FgdHandle *h, int32_t B, float *d_alpha, const float *d_start, const float *d_goal,
                     float *d_fstate, int32_t *d_istate, int32_t max_launch_iters, void *stream);
int fgd_optimize_gd(FgdHandle *h, int32_t B, float *d_alpha, const float *d_start, const float *d_goal,
                    float *d_fstate, int32_t *d_istate, int32_t max_launch_iters, void *stream);

/* The reference's "change the environment at runtime" (README.md:25, blog-post.html:353; plain loop:
 * optimizer_BLS.py:79,82,90 re-reads self.env.obstacles every iteration) WITHOUT leaving the kernel: same operators as
 * fgd_optimize_bls / _gd, run to completion in ONE persistent launch while the host keeps calling
 * fgd_set_obstacles_async on another stream.  Every trajectory team keeps a private copy of the obstacle set in shared
 * memory and polls the "latest generation" word when it picks a trajectory up and then every `poll_every` inner
 * iterations of that trajectory (poll_every >= 1); when the generation changed it copies the new set and re-evaluates
 * the loss at the current alpha before the next gradient step - exactly what a relaunch with
 * max_launch_iters = poll_every does, minus the relaunch, the state write-back / re-fetch and the host round trip.
 * Which generation a trajectory sees at which iteration depends on timing; d_switch_log (optional, may be NULL)
 * records it: int32 [B][FGD_SWITCH_LOG][2], entry 0 = (number of switches n, 0), entries 1..n = (inner iterations
 * completed, generation adopted) - replaying that schedule through the budgeted entry points (or the oracle)
 * reproduces the result bit for bit.  Single-warp teams only (n_timesteps <= 64): FGD_ERR_UNSUPPORTED_T otherwise;
 * FGD_ERR_TOO_MANY_OBSTACLES if obstacle_capacity x 16 teams does not fit in shared memory (capacity <= ~1500);
 * FGD_ERR_INVALID_ARGUMENT for poll_every < 1 or a handle created with whole_arm_cost. */
int fgd_optimize_live(FgdHandle *h, int32_t use_gd, int32_t B, float *d_alpha, const float *d_start, const float *d_goal,
                      float *d_fstate, int32_t *d_istate, int32_t poll_every, int32_t *d_switch_log, void *stream);

/* Same operators with HOST buffers (what a reference-side binding would call):
 * H2D of alpha/start/goal, run to completion, D2H of alpha and state, all on
 * `stream`, synchronised before returning.  use_gd: 0 = BLS, 1 = GD. */
int fgd_optimize_host(FgdHandle *h, int32_t use_gd, int32_t B, float *h_alpha, const float *h_start,
                      const float *h_goal, float *h_fstate, int32_t *h_istate, void *stream);

/* Fresh run with separate input and output host buffers (the shape of the reference's operator: alpha in,
 * new alpha out, `main.py:122`): h_alpha_in is only read, the loop state starts from zero on the device
 * (no state upload) and the final alpha / state are written to the *_out buffers.  Synchronous.
 * When all six buffers are page-locked host memory (cudaHostAlloc, cudaHostRegister, torch pin_memory) the call is
 * ZERO-COPY: the kernel reads each trajectory over PCIe when a team picks it up and writes the result when the team
 * retires it, so the transfers overlap the optimisation of the other trajectories (fgd_zero_copy_calls counts these
 * calls).  Pageable buffers take the staged path (H2D copy, launch, D2H copy); the results are identical. */
int fgd_optimize_host_io(FgdHandle *h, int32_t use_gd, int32_t B, const float *h_alpha_in, float *h_alpha_out,
                         const float *h_start, const float *h_goal, float *h_fstate_out, int32_t *h_istate_out,
                         void *stream);

/* Random-restart reduction: the local trajectories are laid out [n_problems][n_restarts];
 * for each problem pick the restart with the lowest obstacle cost among the
 * constraint-fulfilling ones (falls back to lowest cost if none is fulfilled; ties: lowest index).
 *   global index of local trajectory (p, r) = index_offset + p * problem_stride + r
 *     (problem_stride <= 0: n_restarts, i.e. a contiguous shard of whole problems; a shard that
 *      holds restarts [r0, r0 + n_restarts) of EVERY problem of a sweep with R restarts passes
 *      index_offset = r0, problem_stride = R)
 *   d_best_cost [n_problems], d_best_index [n_problems]: may be NULL
 *   d_best_key  [n_problems] (may be NULL): the order key as one non-negative int64,
 *      (unfulfilled << 62) | (cost bits << 31) | global index   (cost >= 0, index < 2^31),
 *      so that the per-problem winner over several shards is the elementwise MIN of their keys -
 *      the payload of the sweep's single collective (FGD_KEY_* decode it). */
int fgd_argmin_per_problem(FgdHandle *h, int32_t n_problems, int32_t n_restarts, const float *d_fstate,
                           const int32_t *d_istate, int32_t index_offset, int32_t problem_stride, float *d_best_cost,
                           int32_t *d_best_index, int64_t *d_best_key, void *stream);
#define FGD_KEY_INDEX(key) ((int32_t)((key) & 0x7fffffff))
#define FGD_KEY_FULFILLED(key) ((int32_t)((((key) >> 62) & 1) ^ 1))
#define FGD_KEY_COST_BITS(key) ((uint32_t)(((key) >> 31) & 0x7fffffff))     /* IEEE-754 bits of the (non-negative) cost */

/* Trajectory.initTrajectory (trajectory.py:73-78) on the device, for sweeps whose
 * start/goal already live in HBM.  The reference solves
 * K alpha = line J^-1 per trajectory with line = start + c(t) (goal - start); the
 * right-hand side has rank 2 in t, so
 *     alpha[b][t][:] = u[t] * (start_b J^-1) + w[t] * ((goal_b - start_b) J^-1),
 * with u = K^-1 1 and w = K^-1 c solved ONCE on the host (same FP32 LU) and
 * passed in here.  Equal to the reference's init in exact arithmetic; in FP32 both
 * are within the LU residual of the same line (K is numerically singular), so
 * this is offered BESIDE the host init, never instead of it.
 *   h_u, h_w [T], h_jinv [9] row-major: host arrays, copied.
 *   fgd_init_trajectory: d_start/d_goal [B][3] -> d_alpha [B][T][3]. */
int fgd_set_init_basis(FgdHandle *h, const float *h_u, const float *h_w, const float *h_jinv);
int fgd_init_trajectory(FgdHandle *h, int32_t B, const float *d_start, const float *d_goal, float *d_alpha, void *stream);

/* Introspection for the harness. */
int fgd_launch_geometry(const FgdHandle *h, int32_t B, int32_t *grid, int32_t *block, int32_t *smem_bytes,
                        int32_t *warps_per_trajectory);
int64_t fgd_kernel_launches(const FgdHandle *h);   /* kernels launched through this handle so far */
int64_t fgd_zero_copy_calls(const FgdHandle *h);   /* fgd_optimize_host_io calls served without staging copies */
/* BLS launches that ran the speculative line search: for batches of at most 2 x SM-count trajectories (T <= 64, end-effector
 * cost) fgd_optimize_bls gives every trajectory a CTA of four warps that evaluate four consecutive Armijo candidates at
 * once and select the first accepting one in the reference's order (optimizer_BLS.py:131-150) - same iterates, bit for
 * bit, in fewer sequential trips.  Environment FGD_SPEC_MAX_BATCH=0 disables it (A/B measurements). */
int64_t fgd_speculative_launches(const FgdHandle *h);
/* FP32 FFMA throughput of the current device (TFLOP/s, best of 5 launches of a
 * pure-FFMA kernel): the measured denominator of the harness's roofline.frac. */
int fgd_measure_fp32_peak(FgdHandle *h, double *tflops_out, void *stream);
/* MUFU.RCP throughput of the current device (10^12 reciprocals per second, best of 5 launches of
 * a pure rcp.approx kernel): the second roofline of the obstacle stage (environment.py:43,57 -
 * one reciprocal per (sample, obstacle) pair), nominal 148 SM x 16 / clk x 1.965 GHz = 4.65. */
int fgd_measure_mufu_peak(FgdHandle *h, double *trcp_out, void *stream);
int fgd_abi_version(void);
/* Host helper (no GPU needed): the smallest float t with sqrtf(t) >= eps.  The kernels evaluate the end-point predicates
 * ||q[0] - start|| < eps_position, ||v[0]|| < eps_velocity, ... (robot.py:90-101) as  squared norm < t : the correctly
 * rounded square root is monotonic, so the two tests agree for every input.  Exported for the CPU tests. */
float fgd_sqrt_threshold(float eps);

#ifdef __cplusplus
}
#endif
#endif /* FGD_B200_H */
