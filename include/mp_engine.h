/* mp_engine.h -- C ABI of the B200 batched Melting Pot substrate engine (libmpengine.so).
 *
 * This is the drop-in boundary for the reference's hot path. In the reference the path sits
 * behind the pybind11 module `dmlab2d` (third-party), bound at
 *   /root/reference/meltingpot/utils/substrates/builder.py:179-187
 *     env_raw = dmlab2d.Lab2d(_DMLAB2D_ROOT, lab2d_settings_dict)
 *     dmlab2d.Environment(env=env_raw, observation_names=..., seed=seed)
 * and is consumed through Lab2dWrapper.{reset,step,observation,...}
 *   /root/reference/meltingpot/utils/substrates/wrappers/base.py:26-84.
 * Each entry point below names the reference call it replaces. Plain pointers and sizes only:
 * no torch / C++ types cross this boundary. All functions return 0 on success or a negative
 * MP_E_* code; mp_last_error() describes the most recent failure on the calling thread.
 * Nothing here ever falls back to a CPU implementation.
 */
#ifndef MP_ENGINE_H_
#define MP_ENGINE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp_engine* mp_handle;

enum {
  MP_OK = 0,
  MP_E_INVALID = -1,     /* bad argument / malformed blob */
  MP_E_UNSUPPORTED = -2, /* substrate family or parameter outside what the kernels implement */
  MP_E_CUDA = -3,        /* CUDA runtime error (message in mp_last_error) */
  MP_E_NO_DEVICE = -4    /* no usable sm_100 device: the engine refuses to run */
};

/* Option flags for mp_create / mp_set_flags. */
enum {
  MP_FLAG_RENDER_WORLD = 1u << 0, /* produce WORLD.RGB (base_simulation.lua:347-362) */
  MP_FLAG_RENDER_PLAYERS = 1u << 1, /* produce {i}.RGB (avatar_library.lua:264-276) */
  MP_FLAG_DEFAULT = 3u,
  /* Diagnostics for tools/render_ceiling.py (the images are wrong or absent while any is set):
   * time the renderer's store path and its compositing separately. */
  MP_FLAG_DEBUG_NO_COMPOSE = 1u << 4,   /* issue the stores without drawing */
  MP_FLAG_DEBUG_NO_STORE = 1u << 5,     /* draw without storing */
  MP_FLAG_DEBUG_REUSE_RECORDS = 1u << 6, /* per-cell pass only for the first env of each team */
  MP_FLAG_DEBUG_NO_FENCE = 1u << 7,     /* skip the generic->async proxy fence */
  MP_FLAG_DEBUG_TOP_SPRITE_ONLY = 1u << 8, /* every cell drawn as its top sprite */
  MP_FLAG_DEBUG_PLAIN_LANE_MAP = 1u << 9,  /* mp_create only: deal cells to lanes in plain order (A/B) */
  MP_FLAG_DEBUG_SCATTER_LANE_MAP = 1u << 10 /* mp_create only: fully conflict-free dealing that scatters a cell's rows over turns (A/B) */
};

/* Device buffers owned by the engine; valid until mp_destroy. Contents are overwritten by the
 * next mp_step/mp_reset on the same handle. B = num_envs, P = players. */
typedef struct mp_buffers {
  int32_t num_envs, num_players;
  int32_t rgb_h, rgb_w;     /* per-player view in pixels (88 x 88 for clean_up) */
  int32_t world_h, world_w; /* WORLD.RGB in pixels */
  int32_t num_actions;      /* discrete actions per player */
  int32_t num_scalar_obs;   /* per-player f64 observations besides REWARD */
  uint8_t* rgb;             /* u8  [B][P][rgb_h][rgb_w][3]      "{i}.RGB" */
  uint8_t* world_rgb;       /* u8  [B][world_h][world_w][3]     "WORLD.RGB" */
  double* reward;           /* f64 [B][P]                       "{i}.REWARD" */
  double* discount;         /* f64 [B]   0.0 on FIRST/LAST, 1.0 mid-episode */
  int64_t* step_type;       /* i64 [B]   dm_env.StepType: 0 FIRST, 1 MID, 2 LAST */
  double* scalar_obs;       /* f64 [num_scalar_obs][B][P], order of blob section "scalar_obs" */
  int32_t* avatar_state;    /* i32 [B][P][4] x, y, orientation, alive (debug / parity) */
  uint16_t* grid;           /* u16 [B][L][cells_padded] sprite grid (debug / parity) */
  int32_t grid_layers, grid_cells, grid_cells_padded;
  double* timestep_packed;  /* f64 [B][P+2]: reward[0..P), discount, step type -- one buffer for the per-step all-gather */
  /* Events of the current step (the events:add calls of avatar_library.lua:661,1070,1088,
   * component_library.lua:996, clean_up/components.lua:152,402, territory/components.lua:133,168, coins/components.lua:133). Row = (type, a, b)
   * with 1-based player indices: 1 zap(source, target), 2 edible_consumed(player), 3 player_cleaned(player),
   * 4 claimed_resource(player), 5 destroyed_resource(player), 6 sanctioning(source, target),
   * 7 removal_due_to_sanctioning(source, target), 8 coin_consumed(player, 1 if the coin matched the player's type else 0;
   * coins/components.lua:133-137), 9 mining(player, ore type), 10 extraction(player, ore type),
   * 11 extraction_pair(player_a, player_b | ore type << 8) (coop_mining/components.lua:203-234). Rows of one step are in no particular order. max_events is the
   * family's worst case for one step (per avatar: three events per beam-footprint cell + contact events), so
   * event_count never exceeds it and no event is dropped. */
  int32_t* events;          /* i32 [B][max_events][3] */
  int32_t* event_count;     /* i32 [B] */
  int32_t max_events;
  /* reward | discount | step_type | scalar_obs above are carved, in this order, from ONE device allocation of
   * scalar_block_bytes bytes starting at scalar_block (every element is 8 bytes), so a host consumer can fetch all
   * scalar outputs of a step with a single copy (mp_host_outputs.scalar_block). */
  void* scalar_block;
  uint64_t scalar_block_bytes;
  /* After mp_exchange_create: f64 [2][gathered_world * B][P + 2], the timestep_packed rows of EVERY rank's envs in
   * global env order (rank r's envs at rows [r * B, (r + 1) * B)); slot (step & 1) holds the most recent step. */
  double* gathered;
  int32_t gathered_world;
  /* After mp_gather_obs_create: the stacked observations of EVERY rank's envs in global env order, two slots of
   * gathered_obs_slot_bytes bytes each (slot = parity of the render sequence number, mp_gather_obs_slot):
   * gathered_rgb u8 [world * B][P][rgb_h][rgb_w][3] and gathered_world_rgb u8 [world * B][world_h][world_w][3] of slot
   * 0; add gathered_obs_slot_bytes for slot 1. */
  uint8_t* gathered_rgb;
  uint8_t* gathered_world_rgb;
  uint64_t gathered_obs_slot_bytes;
} mp_buffers;

/* Replaces dmlab2d.Lab2d(...) + dmlab2d.Environment(...) (builder.py:182-187) for `num_envs`
 * independent instances on CUDA device `device`. `blob` is a compiled substrate
 * (include/mpb_format.h). Env b uses RNG key `seed + env_index_base + b`, so results do not
 * depend on how envs are sharded over GPUs. Does NOT start an episode; call mp_reset. */
int mp_create(const void* blob, size_t blob_bytes, int num_envs, int device, uint64_t seed,
              uint64_t env_index_base, uint32_t flags, mp_handle* out);

/* Replaces Lab2dWrapper.close (wrappers/base.py:82-84). */
int mp_destroy(mp_handle h);

int mp_set_flags(mp_handle h, uint32_t flags);

/* Replaces dmlab2d.Environment.reset (wrappers/base.py:30-32; api_factory.lua:85-102): every
 * env (or only those with env_mask[b] != 0; `env_mask` is a DEVICE pointer or NULL) starts its
 * next episode and its FIRST observation is rendered. Asynchronous on `stream` (a cudaStream_t). */
int mp_reset(mp_handle h, const uint8_t* env_mask, void* stream);

/* Replaces dmlab2d.Environment.step (wrappers/base.py:34-36; api_factory.lua:104-111) including
 * the DiscreteActionWrapper table lookup (discrete_action_wrapper.py:97-100).
 * `actions` is a DEVICE pointer to int32 [B][P] discrete action ids. Envs whose previous step was
 * LAST ignore the action and start a new episode (FIRST). Runs the state transition and renders
 * all observations. Asynchronous on `stream`. */
int mp_step(mp_handle h, const int32_t* actions, void* stream);

/* State transition only / rendering only (mp_step == mp_step_state + mp_render). */
int mp_step_state(mp_handle h, const int32_t* actions, void* stream);
int mp_render(mp_handle h, void* stream);

/* Replaces Lab2dWrapper.observation / *_spec (wrappers/base.py:38-80): where the outputs live. */
int mp_get_buffers(mp_handle h, mp_buffers* out);

/* Host-buffer convenience used for the end-to-end metric: copies `actions_host` (int32 [B][P],
 * ideally pinned) to the device, steps, renders, copies the requested outputs into the given
 * HOST buffers (any may be NULL to skip) and synchronises `stream`. */
typedef struct mp_host_outputs {
  uint8_t* rgb;
  uint8_t* world_rgb;
  double* reward;
  double* discount;
  int64_t* step_type;
  double* scalar_obs;
  void* scalar_block; /* if non-NULL: receives mp_buffers.scalar_block (scalar_block_bytes bytes) in one transfer and the
                         four scalar pointers above are ignored */
  int32_t* events;      /* i32 [B][max_events][3], or NULL */
  int32_t* event_count; /* i32 [B], or NULL */
} mp_host_outputs;
int mp_step_host(mp_handle h, const int32_t* actions_host, const mp_host_outputs* out, void* stream);
int mp_reset_host(mp_handle h, const mp_host_outputs* out, void* stream);

/* Pipelined form of mp_step_host for host consumers that alternate two output buffer sets (the usual double-buffered
 * actor loop): enqueues H2D(actions) -> state transition -> rendering on `stream` and the device->host copies of
 * the step's outputs on an internal copy stream, then returns WITHOUT synchronising. `slot` (0 or 1) names the
 * device-side image / scalar staging set used; consecutive calls alternate slots, so step t+1's kernels run while
 * step t's observations are still crossing PCIe. mp_wait(h, slot) blocks until the outputs of the last call on
 * that slot are complete in the host buffers. `actions_host` and `out`'s buffers must stay untouched from the call
 * until mp_wait on the same slot returns. Steps are still applied in call order (one state per env). */
int mp_step_host_async(mp_handle h, const int32_t* actions_host, const mp_host_outputs* out, int slot, void* stream);
int mp_wait(mp_handle h, int slot);

/* Stacked timestep across GPUs (the path's only exchange). Envs shard over ranks with no
 * data-path collective; what every rank needs back is ONE stacked [world * B] tensor of reward / discount / step type.
 * Instead of a collective kernel per step, the kernel that follows a state transition in the stream (the renderer,
 * in its prologue; a small delivery kernel when no render follows) writes this rank's rows straight into every
 * rank's `gathered` buffer through NVLink peer mappings (P + 2 remote stores per env and rank, hidden behind the
 * rendering), with no fence on any hot kernel. mp_exchange_wait is the collective point: every rank enqueues it after
 * its step; its one-warp kernel (which fits beside the persistent renderer) first tells the other ranks that this
 * rank's rows are complete -- true by the kernel boundary -- and then waits for theirs. Read mp_buffers.gathered after it.
 *   mp_exchange_create   allocates this rank's exchange block: 256 bytes of per-rank flags followed by `gathered`
 *                        (one allocation, so one IPC handle shares it); returns its device pointer and size;
 *   mp_ipc_export/open   turn a device pointer into a 64-byte CUDA IPC handle + offset inside the driver allocation
 *                        and back (one process per GPU: exchange handle and offset through torch.distributed);
 *   mp_enable_peer_access for ranks that live in ONE process (tests): plain cudaDeviceEnablePeerAccess;
 *   mp_exchange_connect  takes, in rank order, every rank's block as mapped into this process (its own entry = the
 *                        pointer mp_exchange_create returned); from then on every mp_step / mp_step_state /
 *                        mp_reset publishes. All ranks must issue the same sequence of steps and resets (the slot
 *                        is the parity of the launch sequence number);
 *   mp_exchange_wait     enqueues on `stream` (ordered after the step's kernels) the publish-and-wait of the most recent step;
 *                        every rank must call it once per step;
 *   mp_exchange_slot     which half of `gathered` the most recent step was written to (and its sequence number). */
int mp_exchange_create(mp_handle h, int rank, int world, void** block, uint64_t* block_bytes);
int mp_ipc_export(const void* device_ptr, void* handle64, uint64_t* offset);
int mp_ipc_open(int device, const void* handle64, uint64_t offset, void** device_ptr);
int mp_enable_peer_access(int device, int peer_device);
int mp_exchange_connect(mp_handle h, void* const* peer_blocks);
int mp_exchange_wait(mp_handle h, void* stream);
int mp_exchange_slot(mp_handle h, int* slot, uint64_t* step);

/* Stacked OBSERVATIONS across GPUs -- the all-gather BASELINE.json's north_star names ("an NCCL all-gather over NVLink
 * only to return a single stacked observation tensor"), fused into the renderer: with gathering on, k_render hands
 * every finished strip in its shared-memory staging buffer to one TMA bulk store per rank (cp.async.bulk over NVLink
 * peer mappings) in addition to the local one, so the pixels are composed once and travel while the next strips are
 * being drawn -- no collective kernel, no second pass over HBM. NVLink-bound by construction: every rank receives
 * (world - 1) x its own observation bytes per step. Same create / IPC / connect / wait / slot protocol as mp_exchange_*;
 * mp_gather_obs_enable switches the extra stores on and off (they start on after connect). */
int mp_gather_obs_create(mp_handle h, int rank, int world, void** block, uint64_t* block_bytes);
int mp_gather_obs_connect(mp_handle h, void* const* peer_blocks);
int mp_gather_obs_enable(mp_handle h, int on);
int mp_gather_obs_wait(mp_handle h, void* stream);
int mp_gather_obs_slot(mp_handle h, int* slot, uint64_t* step);

/* Number of kernels this engine has launched since creation (all streams). */
int mp_launch_count(mp_handle h, uint64_t* out);

/* Algorithmic bytes one env-step moves (SURVEY.md section 8d formula), for roofline reports. */
int mp_algorithmic_bytes(mp_handle h, uint64_t* per_env_step, uint64_t* render_per_env_step);

/* Snapshot / restore of every env instance. The reference has no counterpart
 * (a dmlab2d env cannot be cloned); here the state is a handful of SoA arrays and the random numbers are
 * addressed by (seed, env, frame), so a byte copy is a complete checkpoint. A snapshot is an opaque string of
 * mp_state_size() bytes in host memory, valid for an engine created from the same blob with the same num_envs,
 * seed and env_index_base: the header records the env count, payload size, RNG key and a hash of the blob, and
 * mp_state_load rejects (MP_E_INVALID) a buffer whose `nbytes` or header does not match this engine.
 * mp_state_load re-renders the observations; both calls synchronise `stream`. */
int mp_state_size(mp_handle h, uint64_t* bytes);
int mp_state_save(mp_handle h, void* host_dst, void* stream);
int mp_state_load(mp_handle h, const void* host_src, uint64_t nbytes, void* stream);

/* Diagnostic: how the renderer was laid out for this substrate: teams per CTA, threads per team, log2 of the pixel
 * rows per WORLD.RGB strip, shared memory bytes, atlas sprites, record stride (u16), staging bytes per warp, grid bytes. */
int mp_debug_render_plan(mp_handle h, int32_t out[8]);

/* Diagnostic: the renderer's sprite tables. *n_total = atlas sprites including the pre-merged ones;
 * pair[n_total * n_total] = pre-merged sprite for (bottom, top) or 0; flags[n_total] bit 0 opaque,
 * bit 1 remapped per viewer. Either array may be NULL (call once for n_total, then again). */
int mp_debug_render_tables(mp_handle h, int32_t* n_total, uint8_t* pair, uint8_t* flags);

/* Debug observations of the current timestep, written into caller-owned DEVICE buffers (any
 * may be NULL): position i32 [B][P][2] = (x, y) of each avatar, 0-based ("{i}.POSITION", LocationObserver,
 * component_library.lua:806-855; an avatar that is off the map keeps its last cell), orientation i32 [B][P] = 0 N, 1 E,
 * 2 S, 3 W ("{i}.ORIENTATION"), layer i32 [B][P][view_h][view_w][L] = the avatar's UNROTATED view window as per-layer
 * sprite ids + 1, 0 = empty, -1 = outside a BOUNDED map ("{i}.LAYER", avatar_library.lua:247-257 with orientation 'N'),
 * zap_matrix i32 [B][P][P] = how often player row zapped player column this step (from the step's 'zap' events, as
 * clean_up.py:751-784's metric does). Off by default in the reference's configs and off the hot path here. */
int mp_debug_observations(mp_handle h, int32_t* position, int32_t* orientation, int32_t* layer, int32_t* zap_matrix, void* stream);

/* Diagnostic (needs no device): the renderer's lane -> cell dealing for a strip of `n_rows` pixel rows x `n_cells` cells
 * at a row pitch of `pitch_slots` 8-byte slots, `iters` turns per lane: out[lane] holds 6 bits per turn (63 = idle).
 * Lane l draws pixel row l % n_rows. scattered = 0: the default dealing (whole cells per lane group and turn, cell order
 * chosen to minimise bank conflicts of the 64-bit staging stores); 1: the fully conflict-free colouring (A/B flag). */
int mp_debug_lane_map(int n_rows, int n_cells, int pitch_slots, int iters, int scattered, uint32_t out[32]);

const char* mp_last_error(void);
const char* mp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MP_ENGINE_H_ */
