/*
 * rnad_b200.h - C ABI of the B200-native R-NaD self-play hot path.
 *
 * The reference (baskuit/R-NaD) has no FFI of its own: its hot path is Python
 * calling ATen.  Each entry point below replaces the group of ATen calls the
 * cited reference lines issue, and is what a reference-side binding (ctypes,
 * see INTEGRATION.md) loads from librnad_b200.so.
 *
 * Conventions (all entry points)
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the
 *     parameter says "host"; nothing is allocated, freed or retained;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL =
 *     the legacy default stream) and the call returns without synchronising;
 *   - the return value is 0 on success, a negative RNAD_E* code otherwise;
 *     rnad_last_error() gives the message for the calling thread;
 *   - time-major trajectory tensors are laid out (T, B, ...) contiguous,
 *     exactly like the reference's `Episodes` members (episode.py:218-227).
 */
#ifndef RNAD_B200_H
#define RNAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RNAD_API __attribute__((visibility("default")))
#else
#define RNAD_API
#endif

#define RNAD_OK 0
#define RNAD_EINVAL (-1)      /* bad argument (shape, null pointer, unsupported size) */
#define RNAD_ECUDA (-2)       /* CUDA runtime / launch error */
#define RNAD_EUNSUPPORTED (-3) /* valid request the selected kernel variant cannot serve */

/* precision / engine selector of the fused rollout */
#define RNAD_PREC_FP32 0 /* FFMA on the CUDA cores, fp32 throughout (validation build) */
#define RNAD_PREC_TF32 1 /* first layer on tcgen05 tensor cores, kind::tf32, fp32 accumulate in TMEM */
#define RNAD_PREC_TF32X2 2 /* both layers on tcgen05 (kind::tf32); the second reads relu(hidden) from tensor memory */
#define RNAD_PREC_F16X2 3 /* the same pipeline with kind::f16: fp16 operands (11-bit significand like tf32, round to nearest
                           * even; observations, weights and activations of this net are far inside fp16's range), fp32
                           * accumulate; K = 16 per MMA: half the tensor-core dispatches, hidden activations packed in pairs */
/* Reproducibility: for one seed, tables and weights every engine writes the same bits on every launch (no result depends
 * on the order in which warps or CTAs were scheduled). */

#define RNAD_MAX_ACTIONS 8
#define RNAD_MAX_TRANSITIONS 8

RNAD_API const char* rnad_last_error(void);
/* version of this ABI (major*100 + minor); 200: rnad_rollout takes seed_dev / stats, learner-step entry points */
RNAD_API int rnad_version(void);
/* number of SMs of the current device (host value), <0 on error */
RNAD_API int rnad_device_sm_count(void);

/* ------------------------------------------------------------------------
 * Packed node tables (replaces the 5 per-node tensors of tree.py:125-140 as
 * kernel input; built once per tree by Tree.packed()).
 *
 *   ev_tab : S x ev_stride 32-bit words.  words [0, A*A) = expected_value[s,0,r,c]
 *            (f32, row-major), word A*A = rows | cols << 8 where the legal mask
 *            of node s is the prefix rectangle r < rows, c < cols (tree.py:133).
 *   tr_tab : S x A*A x tr_stride words, entry (s, r, c) =
 *            [chance[s,0..C-1,r,c] f32 | index[s,0..C-1,r,c] as int32 | value[s,0..C-1,r,c] f32 | pad]
 *   ev_stride = round_up(A*A + 1, 4), tr_stride = round_up(3*C, 4)  (16-byte rows).
 * ------------------------------------------------------------------------ */
RNAD_API int rnad_packed_strides(int A, int C, int* ev_stride, int* tr_stride);

/* index (S,C,A,A) i64, value/chance (S,C,A,A) f32, expected_value/legal (S,1,A,A) f32.
 * bad_flag (device int32, caller zeroes it): set to 1 if some legal mask is not a prefix
 * rectangle, 2 if a child id does not fit int32 or is out of [0,S). */
RNAD_API int rnad_tree_pack(const int64_t* index, const float* value, const float* chance,
                   const float* expected_value, const float* legal,
                   int64_t S, int C, int A,
                   uint32_t* ev_tab, uint32_t* tr_tab, int32_t* bad_flag, void* stream);

/* ------------------------------------------------------------------------
 * K1a  States.observations (episode.py:46-68): obs (B,2,A,A) f32 for the side
 * to move (`turn` 0 = row, 1 = column; all games share it, episode.py:96-98)
 * and the mover's legal-action mask obs[:,1,:,0] (B,A) f32 (may be NULL).
 * S = number of nodes in ev_tab; idx (B) int32 node ids in [0, S).
 * ------------------------------------------------------------------------ */
RNAD_API int rnad_observe(const uint32_t* ev_tab, int A, int64_t S, const int32_t* idx, int turn, int64_t B,
                 float* obs, float* mask, void* stream);

/* ------------------------------------------------------------------------
 * K1b  the transition half of States.step (episode.py:102-124).
 * idx (B) int32 in/out; row_actions/col_actions (B) int64 in [0,A).
 * Chance draw: inverse CDF of chance[s,:,r,c] at u, u = u_chance[b] if
 * u_chance != NULL else Philox4x32-10(key=seed, ctr=(game_offset+b, t, 0)) word 1.
 * reward (B) f32 = value[s,k,r,c] * (s' == 0).  alive (device int32, caller
 * zeroes): incremented by the number of games with s' != 0 (episode.py:124).
 * ------------------------------------------------------------------------ */
RNAD_API int rnad_step(const uint32_t* tr_tab, int A, int C, int32_t* idx,
              const int64_t* row_actions, const int64_t* col_actions,
              const float* u_chance, uint64_t seed, int t, int64_t game_offset,
              int64_t B, float* reward, int32_t* alive, void* stream);

/* Categorical draw replacing torch.multinomial(p, 1) (net.py:49): p (B,N) f32,
 * out (B) int64; u = u_in[b] if u_in != NULL else Philox word 0 of (seed, game, t). */
RNAD_API int rnad_sample_categorical(const float* p, int64_t B, int N, const float* u_in,
                            uint64_t seed, int t, int64_t game_offset, int64_t* out, void* stream);

/* ------------------------------------------------------------------------
 * K2  Episodes.generate (episode.py:175-230) fused with MLP.forward
 * (net.py:37-51): B games from the root for exactly T half-moves.
 * Weights are the reference nn.Linear tensors, fp32, row-major:
 *   value_fc0.weight (W,2A^2) .bias (W) ; value_fc1.weight (1,W) .bias (1)
 *   policy_fc0.weight (W,2A^2) .bias (W); policy_fc1.weight (A,W) .bias (A)
 * uniforms: NULL, or (T,B,2) f32 with [..,0] the action and [..,1] the chance
 * uniform (parity / replay mode).
 * Outputs, all (T,B,...) contiguous, reference dtypes (episode.py:218-225):
 *   indices i64, turns i64, observations f32 (T,B,2,A,A), policy f32 (T,B,A),
 *   actions f32 one-hot (T,B,A), rewards f32, values f32, masks f32 (T,B,A).
 * seed_dev: NULL, or a device uint64 the kernel reads the seed from instead of `seed` (a learner step
 * captured in a CUDA graph varies the seed without re-recording the launch; see rnad_step_control).
 * stats (device int32[4]; the call zeroes it in stream order - also for an empty batch):
 *   [0] the longest game in half-moves = 1 + the last half-move at which some game was not yet on the
 *       absorbing node (t_eff + 1);  [1], [2] the number of valid (t,b) slots of player 0 / player 1 - the
 *       normalisers N_0, N_1 of both losses (vtrace.py:370-374, 387-389), so that nobody has to count them
 *       again;  [3] reserved.
 * workspace: 16-byte aligned device scratch of rnad_rollout_workspace_bytes() (RNAD_PREC_TF32 only: the
 * weight image in MMA operand order; RNAD_PREC_TF32X2, RNAD_PREC_F16X2 and RNAD_PREC_FP32 need none); may be NULL when 0.
 * ------------------------------------------------------------------------ */
typedef struct rnad_mlp_weights {
    const float* value_fc0_w; const float* value_fc0_b;
    const float* value_fc1_w; const float* value_fc1_b;
    const float* policy_fc0_w; const float* policy_fc0_b;
    const float* policy_fc1_w; const float* policy_fc1_b;
    int width;
} rnad_mlp_weights;

typedef struct rnad_trajectory {
    int64_t* indices; int64_t* turns; float* observations; float* policy;
    float* actions; float* rewards; float* values; float* masks;
    /* optional, may be NULL: the policy head's logits (T,B,A) f32 - not part of the reference's Episodes; an on-policy
     * learner (actor == learner net) reuses them and `values` instead of evaluating its own net again (rnad.py:373) */
    float* logits;
    /* optional, may be NULL: every game's payoff for the row player (B) f32 = the sum of its rewards over time */
    float* returns;
} rnad_trajectory;

RNAD_API int rnad_rollout(const uint32_t* ev_tab, const uint32_t* tr_tab, int A, int C,
                 const rnad_mlp_weights* w /* host struct of device pointers */,
                 int64_t B, int T, uint64_t seed, const uint64_t* seed_dev, int64_t game_offset,
                 const float* uniforms, int precision,
                 const rnad_trajectory* out /* host struct of device pointers */,
                 int32_t* stats, void* workspace, void* stream);

RNAD_API int64_t rnad_rollout_workspace_bytes(int A, int width, int precision);

/* 1 if the RNAD_PREC_TF32 engine serves this net shape (width == 256, 2 <= A <= 4), else 0 */
RNAD_API int rnad_rollout_tc_supported(int A, int width);
/* 1 if the RNAD_PREC_TF32X2 / RNAD_PREC_F16X2 engine serves this shape (width == 256, 2 <= A <= 4, max_transitions C <= 4), else 0 */
RNAD_API int rnad_rollout_tc2_supported(int A, int width, int C);

/* ------------------------------------------------------------------------
 * K3  learn/vtrace.py + learn/rnad.py:368-425
 * ------------------------------------------------------------------------ */

/* vtrace.process_policy (vtrace.py:24-55) on n_rows = T*B rows of A entries. */
RNAD_API int rnad_process_policy(const float* policy, const float* mask, int64_t n_rows, int A,
                        int n_disc, float eps_threshold, float* out, void* stream);

/* vtrace.v_trace for ONE player, reference signature (vtrace.py:207-352):
 * v (T,B,1), valid (T,B) f32, player_id (T,B) i64, acting/merged policy and
 * merged_log_policy / actions_oh (T,B,A), player_others (T,B,1), reward (T,B).
 * Outputs v_target (T,B,1) f32, has_played (T,B) i64, learning_output (T,B,A). */
RNAD_API int rnad_vtrace(const float* v, const float* valid, const int64_t* player_id,
                const float* acting_policy, const float* merged_policy, const float* merged_log_policy,
                const float* player_others, const float* actions_oh, const float* reward, int player,
                float eta, float lambda_, float c, float rho, float gamma,
                int T, int64_t B, int A,
                float* v_target, int64_t* has_played, float* learning_output, void* stream);

/* Fused learner targets: everything RNaD.__learn computes between the four
 * forward_batch calls and loss.backward() (rnad.py:365-425), both players in
 * one reverse pass over time, plus the analytic gradients of
 *   loss = value_weight*loss_v + neurd_weight*loss_nerd
 * with respect to the learner net's `logit` and `v` outputs.
 * Inputs (T,B,...): indices i64 (valid = indices != 0), turns i64, mu = the
 * acting policy, actions_oh, rewards (player 0's; player 1 gets the negation),
 * masks, and from the nets: logit/pi/log_pi/v (learner), v_target_net
 * (target), log_pi_reg, log_pi_reg_ (regularisation nets).  pi and log_pi may both be
 * NULL: they are then derived from logit and masks with net.py:76-80's formulas
 * (e = mask ? exp(logit) : 0; pi = e / max(sum e, 1e-12); log_pi = mask ? logit - log(sum e) : 0).
 * Outputs: d_logit (T,B,A), d_v (T,B) required; the rest may be NULL:
 * pi_processed (T,B,A), v_target[2] (T,B), has_played[2] (T,B) i64,
 * learning_output[2] (T,B,A).  losses: device float[2] = {loss_v, loss_nerd};
 * counts: device int32[2] = {N_0, N_1} (sum of has_played).
 * workspace: 16-byte aligned device scratch of rnad_learner_targets_workspace(T,B) bytes, ZEROED once by the caller
 * before its first use (it holds the ticket by which the last block of a launch knows it is the last and adds the
 * per-block loss sums in block order; every launch leaves it zeroed again). */
typedef struct rnad_learner_io {
    const int64_t* indices; const int64_t* turns; const float* mu; const float* actions_oh;
    const float* rewards; const float* masks;
    const float* logit; const float* pi; const float* log_pi; const float* v;
    const float* v_target_net; const float* log_pi_reg; const float* log_pi_reg_;
    float* d_logit; float* d_v;
    float* pi_processed; float* v_target[2]; int64_t* has_played[2]; float* learning_output[2];
    float* losses; int32_t* counts;
    /* exact data-parallel normalisation (DP note i): if non-NULL, device int32[2]
     * GLOBAL counts to divide by instead of the local ones (already all-reduced by the caller) */
    const int32_t* global_counts;
    /* unnormalised mode (the captured learner step, csrc/learner_step.cu): d_logit / d_v are written as if N_0 = N_1 = 1
     * and the four loss numerators (critic p0, p1, NeuRD p0, p1) go to loss_sums (device float[4]); counts, losses and
     * global_counts are not touched and nothing is counted - the division by the (global) step counts happens after
     * the gradient exchange.  Rows of even t must belong to player 0 and of odd t to player 1 (rnad_rollout's). */
    int unnormalised;
    float* loss_sums;
} rnad_learner_io;

typedef struct rnad_learner_params {
    float alpha, eta, lambda_, c, rho, gamma;
    float eps_threshold; int n_disc;
    float neurd_clip, beta;
    float value_weight, neurd_weight;
    const float* alpha_dev;   /* if non-NULL: alpha is read from this device float (a captured step varies it) */
} rnad_learner_params;

RNAD_API int64_t rnad_learner_targets_workspace(int T, int64_t B);
/* phase 1 only: counts[p] = sum over (t,b) of (indices != 0 && turns == p) */
RNAD_API int rnad_count_played(const int64_t* indices, const int64_t* turns, int T, int64_t B,
                      int32_t* counts, void* stream);
RNAD_API int rnad_learner_targets(const rnad_learner_io* io /* host struct */, const rnad_learner_params* p,
                         int T, int64_t B, int A, void* workspace, void* stream);

/* ------------------------------------------------------------------------
 * Learner-side net passes of RNaD.__learn, fused (rnad.py:373-380, 424-425;
 * net.py:64-85 forward_batch).  observations: (N, 2A^2) f32 = the (T,B,2,A,A)
 * trajectory tensor flattened, N = T*B.  Tensor-core engine only: width 256,
 * 2 <= A <= 4 (rnad_learner_mlp_supported); otherwise the caller keeps the
 * reference-style batched GEMM path.
 *
 * rnad_learner_forward: ONE pass computes, per row,
 *   learner net : logit (N,A), pi (N,A), log_pi (N,A), v (N)      [forward_batch outputs]
 *   target net  : v_target (N)                                     [value trunk only]
 *   reg nets    : log_pi_reg (N,A), log_pi_reg_ (N,A)              [policy trunks only]
 * rnad_learner_backward: parameter gradients of the learner net for given
 *   d_logit (N,A), d_v (N) (from rnad_learner_targets), written (not accumulated)
 *   to flat_grad, rnad_learner_param_count() floats in state_dict order:
 *   value_fc0.weight, .bias, value_fc1.weight, .bias, policy_fc0.weight, .bias,
 *   policy_fc1.weight, .bias.  Deterministic (fixed-order reduction).
 * workspace: 256-byte aligned device scratch of rnad_learner_mlp_workspace_bytes().
 * ------------------------------------------------------------------------ */
typedef struct rnad_learner_fwd_out {
    float* logit; float* pi; float* log_pi; float* v;
    float* v_target; float* log_pi_reg; float* log_pi_reg_;
} rnad_learner_fwd_out;

RNAD_API int rnad_learner_mlp_supported(int A, int width);
RNAD_API int64_t rnad_learner_mlp_workspace_bytes(int A, int width);
RNAD_API int rnad_learner_param_count(int A, int width);
RNAD_API int rnad_learner_forward(const float* observations, int64_t N, int A,
                         const rnad_mlp_weights* net, const rnad_mlp_weights* target,
                         const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_,
                         const rnad_learner_fwd_out* out, void* workspace, void* stream);
RNAD_API int rnad_learner_backward(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                          const float* d_logit, const float* d_v, float* flat_grad,
                          void* workspace, void* stream);
/* The same for a (T,B,...) trajectory, one gradient per PLAYER: rows of even t are player 0's steps, rows of odd t
 * player 1's (every row has exactly one owner and d_logit / d_v are zero elsewhere), and player_grads receives
 * 2 x rnad_learner_param_count() floats - player 0's gradient, then player 1's.  For max_actions <= 3 this call
 * runs on the fp16-operand engine (csrc/learner_bwd_f16.cu; fp32 accumulation): d_logit / d_v - and their products
 * with the observations - are rounded to fp16 (11-bit significand like tf32; saturating at +-65504, subnormal below
 * 6e-5), which suits UNNORMALISED gradients; pass gradients that were already divided by large step counts to
 * rnad_learner_backward instead (tf32 operands).  With d_logit / d_v from
 * rnad_learner_targets in unnormalised mode these are the numerators G_p of
 *     d loss / d params = G_0 / N_0 + G_1 / N_1        (vtrace.py:370-374, 387-389; N_p = the players' step counts)
 * which is what data-parallel ranks exchange: the division by the GLOBAL counts happens once, after the sum over
 * ranks (rnad_learner_tail). */
RNAD_API int rnad_learner_backward_split(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                const float* d_logit, const float* d_v, float* player_grads,
                                void* workspace, void* stream);

/* The weight images (MMA operand order) of rnad_learner_forward and rnad_learner_backward(_split) depend only on the
 * nets: rnad_learner_pack writes them into `workspace` (what those calls otherwise do first), and the *_prepacked
 * variants then skip it - so that a captured learner step can pack on a side stream while the rollout runs. */
/* others_only != 0 (in both calls alike): only the target net's value trunk and the regularisation nets' policy
 * trunks are packed / evaluated - v_target, log_pi_reg, log_pi_reg_ are written, the learner's logit / pi / log_pi / v
 * outputs are not touched (an on-policy step takes them from the rollout: rnad_trajectory.logits, .values). */
RNAD_API int rnad_learner_pack(int A, const rnad_mlp_weights* net, const rnad_mlp_weights* target,
                      const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_, int others_only,
                      void* workspace, void* stream);
RNAD_API int rnad_learner_forward_prepacked(const float* observations, int64_t N, int A,
                         const rnad_mlp_weights* net, const rnad_mlp_weights* target,
                         const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_,
                         const rnad_learner_fwd_out* out, int others_only, void* workspace, void* stream);
RNAD_API int rnad_learner_backward_split_prepacked(const float* observations, int T, int64_t B, int A,
                                const rnad_mlp_weights* net, const float* d_logit, const float* d_v,
                                float* player_grads, void* workspace, void* stream);

/* ------------------------------------------------------------------------
 * One learner step as a replayable unit (rnad.py:495-526 loop body): the per-step scalars live in a device
 * control block, the optimizer tail - and under data parallelism the one gradient exchange - is one kernel.
 * ------------------------------------------------------------------------ */
#define RNAD_MAX_PEERS 16
#define RNAD_IPC_HANDLE_BYTES 64

typedef struct rnad_step_ctrl {   /* DEVICE memory, 32 bytes, zero-initialised by the caller */
    uint64_t seed;                /* rollout seed of this step      (pass &ctrl->seed  as rnad_rollout's seed_dev) */
    float alpha;                  /* reward-transform mixing weight (pass &ctrl->alpha as rnad_learner_params.alpha_dev) */
    uint32_t seq;                 /* completed rnad_learner_tail calls: slot parity and flag value of the exchange */
    float adam_step;              /* Adam's step count (torch keeps it as a float) */
    uint32_t error;               /* bit r set: gave up waiting for rank r's gradients */
    uint32_t seed_state[2];       /* rnad_step_advance: 64-bit splitmix64 state (low word, high word) */
} rnad_step_ctrl;

/* writes seed and alpha into *ctrl in stream order (a one-thread kernel; the arguments travel by value) */
RNAD_API int rnad_step_control(rnad_step_ctrl* ctrl, uint64_t seed, float alpha, void* stream);
/* the next rollout seed WITHOUT the host: state += 0x9E3779B97F4A7C15, seed = splitmix64's mix of the state >> 2 (a
 * one-thread kernel; inside a captured graph every replay plays a new, host-predictable seed) */
RNAD_API int rnad_step_advance(rnad_step_ctrl* ctrl, void* stream);
/* rnad_step_advance and, in the same kernel, the step's fresh inputs: n floats from src to dst (both 16-byte aligned).
 * src may be PINNED HOST memory (device-readable at the same address under unified addressing): the actor weights of a
 * self-play batch then arrive with one round trip over PCIe and without a copy node of their own
 * (environment.episode.SelfPlay; the reference's analogue is net.to(device) before Episodes.generate, rnad.py:502). */
RNAD_API int rnad_step_advance_fetch(rnad_step_ctrl* ctrl, const float* src, float* dst, int64_t n, void* stream);

/* rnad_learner_tail: [sum over ranks of (G_0 | G_1 | N_0, N_1 | loss numerators) over NVLink peer memory] ->
 * g = G_0 / N_0 + G_1 / N_1 -> clip_grad_norm_(grad_clip) (rnad.py:456) -> Adam (torch.optim.Adam semantics without
 * amsgrad / weight decay, rnad.py:514; exp_avg, exp_avg_sq and ctrl->adam_step are its state) ->
 * target = gamma_averaging * params + (1 - gamma_averaging) * target (rnad.py:516-523).
 * All flat arrays hold n_params floats in state_dict order (see rnad_learner_backward).
 * flat_grad: the clipped gradient (what p.grad holds after RNaD.__learn).
 * losses: device float[4] = {loss_v, loss_nerd, gradient norm before clipping, ctrl->error}; under data parallelism
 * the losses are those of the GLOBAL batch, identical on every rank.
 * world == 1: no exchange, xchg is ignored.  world > 1: xchg[r] is rank r's exchange buffer as mapped into THIS
 * process (xchg[rank] the local one), each rnad_xchg_bytes(n_params, world) bytes, created zeroed by
 * rnad_xchg_create on its owner and opened by the others through the 64-byte CUDA IPC handle; every rank calls
 * rnad_learner_tail once per step, all with the same ctrl->seq. */
typedef struct rnad_tail_args {
    int n_params;
    const float* player_grads;    /* [2][n_params], this rank's (rnad_learner_backward_split) */
    const int32_t* stats;         /* rnad_rollout's stats words: [1], [2] = this rank's N_0, N_1 */
    const float* loss_sums;       /* float[4], rnad_learner_targets in unnormalised mode */
    float* params; float* target_params; float* exp_avg; float* exp_avg_sq;
    float* flat_grad; float* losses;
    rnad_step_ctrl* ctrl;
    float lr, beta1, beta2, eps, grad_clip, gamma_averaging, one_minus_gamma_averaging;
    int world, rank;
    float* xchg[RNAD_MAX_PEERS];
    float* losses_host;           /* optional: the same four floats once more, e.g. into PINNED HOST memory (device-
                                   * writable at the same address under unified addressing) - a host that wants the
                                   * losses every step then needs a stream synchronize and no copy */
} rnad_tail_args;

RNAD_API int rnad_learner_tail(const rnad_tail_args* args /* host struct */, void* stream);

RNAD_API int64_t rnad_xchg_bytes(int n_params, int world);
/* cudaMalloc + zero + cudaIpcGetMemHandle: *ptr = the local buffer, handle = RNAD_IPC_HANDLE_BYTES bytes to hand to the
 * other ranks (host memory; e.g. through torch.distributed.all_gather_object) */
RNAD_API int rnad_xchg_create(int64_t bytes, void** ptr, unsigned char* handle);
RNAD_API int rnad_xchg_open(const unsigned char* handle, void** ptr);   /* a peer's buffer, mapped into this process */
RNAD_API int rnad_xchg_close(void* ptr);                                 /* unmap a peer's buffer */
RNAD_API int rnad_xchg_destroy(void* ptr);                               /* free the local buffer */

#ifdef __cplusplus
}
#endif
#endif /* RNAD_B200_H */
