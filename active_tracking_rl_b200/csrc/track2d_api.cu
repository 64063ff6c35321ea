// track2d_api.cu -- the C ABI declared in include/track2d.h: handle lifetime, stream-ordered
// reset/step over caller-owned device buffers, the host-buffer variants, and synchronous state
// read-back / injection used by the parity tests and the single-env gym shim.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "track2d_common.cuh"

// kernels (track2d_step.cu / track2d_reset.cu)
cudaError_t t2d_launch_step_f32(const World &w, const int32_t *actions, float *obs, float *reward, uint8_t *done, cudaStream_t s);
cudaError_t t2d_launch_step_u8(const World &w, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done, cudaStream_t s);
cudaError_t t2d_launch_full_obs_f32(const World &w, float *obs, const uint8_t *mask, cudaStream_t s);
cudaError_t t2d_launch_full_obs_u8(const World &w, uint8_t *obs, const uint8_t *mask, cudaStream_t s);
cudaError_t t2d_launch_reset_f32(const World &w, const uint8_t *mask, int from_list, float *obs, int init_only, cudaStream_t s);
cudaError_t t2d_launch_reset_u8(const World &w, const uint8_t *mask, int from_list, uint8_t *obs, int init_only, cudaStream_t s);
cudaError_t t2d_launch_seed_numpy(const World &w, int first, int count, unsigned long long seed, int add_index, cudaStream_t s);
cudaError_t t2d_launch_nav_replan(const World &w, cudaStream_t s);
cudaError_t t2d_launch_astar_direct(const World &w, int first, int count, const int32_t *sg_dev, int32_t *len_dev, cudaStream_t s);
cudaError_t t2d_launch_nav_fill(const World &w, const uint8_t *mask, const uint32_t *list, const uint32_t *count, cudaStream_t s);
cudaError_t t2d_launch_nav_ahead(const World &w, uint32_t *list, uint32_t *count, cudaStream_t s);
cudaError_t t2d_launch_nav_merge(const World &w, cudaStream_t s);
cudaError_t t2d_launch_swap_standby(const World &w, const World &sw, void *obs, int obs_u8, uint32_t *regen_list, uint32_t *regen_count, cudaStream_t s);
int t2d_nav_slots();

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0; // kernels of this library enqueued so far (host-side count)
void t2d_count_launches(int n) { g_launches += (unsigned long long)n; }

void t2d_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct track2d_env {
    track2d_config cfg;
    World w;
    cudaStream_t own_stream;
    // device staging for the host-buffer API
    int32_t *d_actions;
    float *d_obs;
    uint8_t *d_obs8;
    float *d_reward;
    uint8_t *d_done;
    uint8_t *d_mask;
    cudaEvent_t chunk_ev[16];  // track2d_step_host_begin: one event per observation chunk
    int n_chunk_ev;
    // standby worlds / plan-ahead (Philox, auto-reset; Nav targets and Maze maps): see track2d_reset.cu
    bool standby;              // auto-reset = copy of a world prepared ahead on the side stream
    World sw;                  // the standby world (same layout, its own arrays)
    cudaStream_t side;
    cudaEvent_t ev_main[8], ev_side[8];
    unsigned since_join;       // steps issued since the side stream was last joined
    uint32_t *regen_lists, *regen_counts, *ahead_list, *ahead_count;
    uint8_t *side_ws;
    bool was_reset;
    unsigned long long steps_done;
    std::vector<void *> allocs;
};

namespace {

struct DeviceGuard { // run on the handle's device without disturbing the caller's current device
    int prev;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define T2D_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t err__ = (call);                                                     \
        if (err__ != cudaSuccess) {                                                     \
            t2d_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return T2D_E_CUDA;                                                          \
        }                                                                               \
    } while (0)

#define T2D_REQUIRE(cond, msg)               \
    do {                                     \
        if (!(cond)) {                       \
            t2d_set_error("%s", msg);        \
            return T2D_E_INVALID;            \
        }                                    \
    } while (0)

template <typename T>
int dev_alloc(track2d_env *env, T **ptr, size_t count, bool zero = true) {
    void *p = nullptr;
    T2D_CUDA(cudaMalloc(&p, count * sizeof(T)));
    env->allocs.push_back(p);
    if (zero) T2D_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    *ptr = reinterpret_cast<T *>(p);
    return T2D_OK;
}

int check_range(const track2d_env *env, int first, int count) {
    if (!env) { t2d_set_error("null handle"); return T2D_E_INVALID; }
    if (first < 0 || count < 0 || first + count > env->w.E) {
        t2d_set_error("env range [%d, %d) outside [0, %d)", first, first + count, env->w.E);
        return T2D_E_INVALID;
    }
    return T2D_OK;
}

inline int cells_of(const track2d_env *env) { return env->cfg.obs_type == T2D_OBS_FULL ? env->w.H * env->w.W : T2D_WIN_CELLS; }

template <typename ObsT>
int do_reset(track2d_env *env, const uint8_t *mask, ObsT *obs, int init_only, cudaStream_t s) {
    const World &w = env->w;
    cudaError_t err;
    if (sizeof(ObsT) == 4) err = t2d_launch_reset_f32(w, mask, 0, (float *)obs, init_only, s);
    else err = t2d_launch_reset_u8(w, mask, 0, (uint8_t *)obs, init_only, s);
    T2D_CUDA(err);
    t2d_count_launches(1);
    if (!init_only && obs && w.obs_type == T2D_OBS_FULL) {
        if (sizeof(ObsT) == 4) err = t2d_launch_full_obs_f32(w, (float *)obs, mask, s);
        else err = t2d_launch_full_obs_u8(w, (uint8_t *)obs, mask, s);
        T2D_CUDA(err);
    }
    if (!init_only && env->standby) {
        // live look-ahead, then the standby worlds of the envs that were just reset (rare: done on the caller's stream)
        if (w.async_nav) T2D_CUDA(t2d_launch_nav_fill(w, mask, nullptr, nullptr, s));
        T2D_CUDA(cudaMemcpyAsync(env->sw.episode, w.episode, (size_t)w.E * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        T2D_CUDA(t2d_launch_reset_u8(env->sw, mask, 0, nullptr, 0, s));
        if (w.async_nav) T2D_CUDA(t2d_launch_nav_fill(env->sw, mask, nullptr, nullptr, s));
        t2d_count_launches(w.async_nav ? 3 : 1);
    }
    if (!init_only) env->was_reset = true;
    return T2D_OK;
}

template <typename ObsT>
int do_step(track2d_env *env, const int32_t *actions, ObsT *obs, float *reward, uint8_t *done, cudaStream_t s) {
    const World &w = env->w;
    if (!env->was_reset) { // gym TimeLimit asserts "Cannot call env.step() before calling reset()"
        t2d_set_error("step() before reset()");
        return T2D_E_STATE;
    }
    T2D_REQUIRE(actions && reward && done, "step: actions, reward and done buffers are required");
    T2D_REQUIRE(obs == nullptr || ((uintptr_t)obs & 15u) == 0, "step: obs buffer must be 16-byte aligned");
    T2D_REQUIRE(((uintptr_t)actions & 7u) == 0 && ((uintptr_t)reward & 7u) == 0, "step: actions/reward buffers must be 8-byte aligned");
    cudaError_t err;
    if (env->standby) {
        constexpr unsigned R = 8, LAG = 4;
        const unsigned j = env->since_join, slot = j % R;
        if (j >= LAG) T2D_CUDA(cudaStreamWaitEvent(s, env->ev_side[(j - LAG) % R], 0));  // side work issued LAG steps ago is done
        if (w.async_nav) {
            T2D_CUDA(t2d_launch_nav_merge(w, s));  // hand-over of finished segments + the synchronous fallback for a starved target
            t2d_count_launches(4);
        }
        if (sizeof(ObsT) == 4) err = t2d_launch_step_f32(w, actions, (float *)obs, reward, done, s);
        else err = t2d_launch_step_u8(w, actions, (uint8_t *)obs, reward, done, s);
        T2D_CUDA(err);
        uint32_t *rl = env->regen_lists + (size_t)slot * w.E, *rc4 = env->regen_counts + slot * 4;
        T2D_CUDA(cudaMemsetAsync(rc4, 0, 4 * sizeof(uint32_t), s));
        T2D_CUDA(t2d_launch_swap_standby(w, env->sw, obs, sizeof(ObsT) == 1, rl, rc4, s));
        t2d_count_launches(2);
        // side stream: new standby worlds for the envs that were just swapped in, then plan ahead on the live world
        T2D_CUDA(cudaEventRecord(env->ev_main[slot], s));
        T2D_CUDA(cudaStreamWaitEvent(env->side, env->ev_main[slot], 0));
        World sws = env->sw;
        sws.work_list = rl;
        sws.work_count = rc4;
        T2D_CUDA(t2d_launch_reset_u8(sws, nullptr, 1, nullptr, 0, env->side));
        t2d_count_launches(1);
        if (w.async_nav) {
            T2D_CUDA(t2d_launch_nav_fill(sws, nullptr, rl, rc4 + 3, env->side));
            World wl = w;
            wl.astar_ws = env->side_ws;
            T2D_CUDA(t2d_launch_nav_ahead(wl, env->ahead_list, env->ahead_count, env->side));
            t2d_count_launches(3);
        }
        T2D_CUDA(cudaEventRecord(env->ev_side[slot], env->side));
        env->since_join = j + 1;
        env->steps_done += (unsigned long long)w.E;
        return T2D_OK;
    }
    T2D_CUDA(t2d_launch_nav_replan(w, s));
    if (w.target_mode == T2D_TARGET_NAV || w.target_mode == T2D_TARGET_RPF) t2d_count_launches(3);
    if (sizeof(ObsT) == 4) err = t2d_launch_step_f32(w, actions, (float *)obs, reward, done, s);
    else err = t2d_launch_step_u8(w, actions, (uint8_t *)obs, reward, done, s);
    T2D_CUDA(err);
    t2d_count_launches(1 + ((w.flags & T2D_FLAG_AUTO_RESET) ? 1 : 0));
    if (obs && w.obs_type == T2D_OBS_FULL) {
        if (sizeof(ObsT) == 4) err = t2d_launch_full_obs_f32(w, (float *)obs, nullptr, s);
        else err = t2d_launch_full_obs_u8(w, (uint8_t *)obs, nullptr, s);
        T2D_CUDA(err);
    }
    if (w.flags & T2D_FLAG_AUTO_RESET) {
        if (sizeof(ObsT) == 4) err = t2d_launch_reset_f32(w, nullptr, 1, (float *)obs, 0, s);
        else err = t2d_launch_reset_u8(w, nullptr, 1, (uint8_t *)obs, 0, s);
        T2D_CUDA(err);
        if (obs && w.obs_type == T2D_OBS_FULL) { // `done` marks exactly the envs that were just reset
            if (sizeof(ObsT) == 4) err = t2d_launch_full_obs_f32(w, (float *)obs, done, s);
            else err = t2d_launch_full_obs_u8(w, (uint8_t *)obs, done, s);
            T2D_CUDA(err);
        }
    }
    env->steps_done += (unsigned long long)w.E;
    return T2D_OK;
}

// host <-> bit-packed map conversion
void pack_map(const uint8_t *maze, int H, int W, uint32_t *words) {
    for (int i = 0; i < T2D_MAP_WORDS; i++) words[i] = 0xFFFFFFFFu;
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++)
            if (maze[r * W + c] == 0) {
                int pr = r + T2D_PAD, pc = c + T2D_PAD;
                words[pr * T2D_ROW_WORDS + (pc >> 5)] &= ~(1u << (pc & 31));
            }
}
void unpack_map(const uint32_t *words, int H, int W, uint8_t *maze) {
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++) {
            int pr = r + T2D_PAD, pc = c + T2D_PAD;
            maze[r * W + c] = (words[pr * T2D_ROW_WORDS + (pc >> 5)] >> (pc & 31)) & 1u;
        }
}

} // namespace

extern "C" {

const char *track2d_last_error(void) { return g_err; }
uint64_t track2d_launch_count(void) { return g_launches; }
int track2d_abi_version(void) { return TRACK2D_ABI_VERSION; }

int track2d_create(const track2d_config *cfg, track2d_env **out) {
    T2D_REQUIRE(cfg && out, "create: null argument");
    T2D_REQUIRE(cfg->abi_version == TRACK2D_ABI_VERSION, "create: ABI version mismatch");
    T2D_REQUIRE(cfg->num_envs >= 1, "create: num_envs must be >= 1");
    T2D_REQUIRE(cfg->map_type >= 0 && cfg->map_type <= 2, "create: map_type must be Block/Maze/Empty");
    T2D_REQUIRE(cfg->obs_type == T2D_OBS_PARTIAL || cfg->obs_type == T2D_OBS_FULL, "create: obs_type must be 'Full' or 'Partial'");
    T2D_REQUIRE(cfg->target_mode >= 0 && cfg->target_mode <= 5, "create: unknown target_mode");
    T2D_REQUIRE(cfg->rng_mode == T2D_RNG_PHILOX || cfg->rng_mode == T2D_RNG_NUMPY, "create: unknown rng_mode");
    T2D_REQUIRE(cfg->level >= 0 && cfg->level <= 9, "create: level out of range");
    T2D_REQUIRE(cfg->max_episode_steps < 65535, "create: max_episode_steps must be < 65535");
    int ndev = 0;
    T2D_CUDA(cudaGetDeviceCount(&ndev));
    T2D_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "create: no such CUDA device");
    DeviceGuard guard(cfg->device);

    track2d_env *env = new track2d_env();
    env->cfg = *cfg;
    env->was_reset = false;
    env->steps_done = 0;
    env->d_actions = nullptr; env->d_obs = nullptr; env->d_obs8 = nullptr; env->d_reward = nullptr; env->d_done = nullptr; env->d_mask = nullptr;
    World &w = env->w;
    memset(&w, 0, sizeof(w));
    const int E = cfg->num_envs;
    w.E = E;
    w.H = w.W = cfg->map_type == T2D_MAP_MAZE ? 81 : 82;
    w.map_type = cfg->map_type; w.obs_type = cfg->obs_type; w.target_mode = cfg->target_mode; w.level = cfg->level;
    w.rng_mode = cfg->rng_mode; w.max_steps = cfg->max_episode_steps; w.flags = cfg->flags; w.seed = cfg->seed;
    const bool nav = cfg->target_mode == T2D_TARGET_NAV || cfg->target_mode == T2D_TARGET_RPF;

    int rc = T2D_OK;
#define ALLOC(ptr, count) if (rc == T2D_OK) rc = dev_alloc(env, &(ptr), (size_t)(count))
    ALLOC(w.maps, (size_t)E * T2D_MAP_WORDS);
    ALLOC(w.pos, E);
    ALLOC(w.ctr, E);
    ALLOC(w.goals, E);
    ALLOC(w.ram, E);
    ALLOC(w.episode, E);
    ALLOC(w.rpf, E);
    ALLOC(w.tgt_act, E);
    ALLOC(w.work_list, 2 * (size_t)E);
    ALLOC(w.work_count, 4);
    ALLOC(w.status, 1);
    ALLOC(w.stats, 2);
    ALLOC(w.nav_meta, E);
    ALLOC(w.nav_goal, E);
    if (nav) {
        ALLOC(w.nav_plan, (size_t)E * T2D_NAV_PLAN_BYTES);
        w.astar_slots = t2d_nav_slots();
        ALLOC(w.astar_ws, (size_t)w.astar_slots * (82 * 82) * 16);  // HeapEntry spill area per concurrent plan
    }
    if (cfg->rng_mode == T2D_RNG_NUMPY) {
        ALLOC(w.mt_key, (size_t)E * T2D_MT_N);
        ALLOC(w.mt_pos, E);
    }
    if (cfg->flags & T2D_FLAG_KEEP_F64) ALLOC(w.rew64, 2 * (size_t)E);
    // Opt-in (T2D_FLAG_PLAN_AHEAD): auto-resets as copies of worlds prepared ahead of time, Nav plans made ahead of time.  The RNG is
    // counter-based, so WHEN something is computed cannot change WHAT comes out.
    env->standby = cfg->rng_mode == T2D_RNG_PHILOX && (cfg->flags & T2D_FLAG_AUTO_RESET) && (cfg->flags & T2D_FLAG_PLAN_AHEAD) &&
                   cfg->obs_type == T2D_OBS_PARTIAL && (cfg->target_mode == T2D_TARGET_NAV || (cfg->map_type == T2D_MAP_MAZE && cfg->target_mode != T2D_TARGET_RPF));
    if (env->standby) {
        w.async_nav = cfg->target_mode == T2D_TARGET_NAV;
        World &sw = env->sw;
        sw = w;
        ALLOC(sw.maps, (size_t)E * T2D_MAP_WORDS);
        ALLOC(sw.pos, E);
        ALLOC(sw.ctr, E);
        ALLOC(sw.goals, E);
        ALLOC(sw.ram, E);
        ALLOC(sw.episode, E);
        ALLOC(sw.rpf, E);
        ALLOC(sw.status, 1);
        ALLOC(sw.stats, 2);
        ALLOC(sw.nav_meta, E);
        ALLOC(sw.nav_goal, E);
        ALLOC(env->regen_lists, 8 * (size_t)E);
        ALLOC(env->regen_counts, 8 * 4);
        if (w.async_nav) {
            ALLOC(w.nav_end, E);
            ALLOC(w.ext_info, E);
            ALLOC(w.ext_plan, (size_t)E * T2D_NAV_PLAN_BYTES);
            ALLOC(sw.nav_plan, (size_t)E * T2D_NAV_PLAN_BYTES);
            ALLOC(sw.nav_end, E);
            ALLOC(env->side_ws, (size_t)w.astar_slots * (82 * 82) * 16);
            ALLOC(env->ahead_list, E);
            ALLOC(env->ahead_count, 4);
            sw.astar_ws = env->side_ws;
            sw.ext_info = nullptr;
            sw.ext_plan = nullptr;
        }
        sw.rew64 = nullptr;
        sw.flags = cfg->flags;
    }
#undef ALLOC
    if (rc == T2D_OK && env->standby) {
        bool ok = cudaStreamCreateWithFlags(&env->side, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 8 && ok; i++)
            ok = cudaEventCreateWithFlags(&env->ev_main[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&env->ev_side[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            t2d_set_error("create: side stream / events: %s", cudaGetErrorString(cudaGetLastError()));
            rc = T2D_E_CUDA;
        }
    }
    if (rc == T2D_OK && cudaStreamCreateWithFlags(&env->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        t2d_set_error("create: cudaStreamCreate failed");
        rc = T2D_E_CUDA;
    }
    if (rc == T2D_OK && cfg->rng_mode == T2D_RNG_NUMPY) {
        // env e starts like the reference after np.random.seed(seed + e)
        if (t2d_launch_seed_numpy(w, 0, E, cfg->seed, 1, 0) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            t2d_set_error("create: seeding failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = T2D_E_CUDA;
        }
    }
    if (rc != T2D_OK) {
        for (void *p : env->allocs) cudaFree(p);
        delete env;
        return rc;
    }
    *out = env;
    return T2D_OK;
}

int track2d_destroy(track2d_env *env) {
    if (!env) return T2D_OK;
    DeviceGuard guard(env->cfg.device);
    cudaDeviceSynchronize();
    for (void *p : env->allocs) cudaFree(p);
    for (int c = 0; c < env->n_chunk_ev; c++) cudaEventDestroy(env->chunk_ev[c]);
    if (env->standby) {
        for (int i = 0; i < 8; i++) {
            if (env->ev_main[i]) cudaEventDestroy(env->ev_main[i]);
            if (env->ev_side[i]) cudaEventDestroy(env->ev_side[i]);
        }
        if (env->side) cudaStreamDestroy(env->side);
    }
    if (env->own_stream) cudaStreamDestroy(env->own_stream);
    delete env;
    return T2D_OK;
}

int track2d_obs_cells(const track2d_env *env) { return env ? cells_of(env) : T2D_E_INVALID; }
int track2d_num_envs(const track2d_env *env) { return env ? env->w.E : T2D_E_INVALID; }
int track2d_map_height(const track2d_env *env) { return env ? env->w.H : T2D_E_INVALID; }
int track2d_map_width(const track2d_env *env) { return env ? env->w.W : T2D_E_INVALID; }

int track2d_reset(track2d_env *env, const uint8_t *mask_dev, float *obs_dev, void *stream) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return do_reset<float>(env, mask_dev, obs_dev, 0, (cudaStream_t)stream);
}
int track2d_reset_u8(track2d_env *env, const uint8_t *mask_dev, uint8_t *obs_dev, void *stream) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return do_reset<uint8_t>(env, mask_dev, obs_dev, 0, (cudaStream_t)stream);
}
int track2d_init_maze(track2d_env *env, const uint8_t *mask_dev, void *stream) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return do_reset<float>(env, mask_dev, nullptr, 1, (cudaStream_t)stream);
}
int track2d_step(track2d_env *env, const int32_t *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev, void *stream) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return do_step<float>(env, actions_dev, obs_dev, reward_dev, done_dev, (cudaStream_t)stream);
}
int track2d_step_u8(track2d_env *env, const int32_t *actions_dev, uint8_t *obs_dev, float *reward_dev, uint8_t *done_dev, void *stream) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return do_step<uint8_t>(env, actions_dev, obs_dev, reward_dev, done_dev, (cudaStream_t)stream);
}

// ---- host-buffer API -------------------------------------------------------------------------------
} // extern "C" (the templated helpers below need C++ linkage)

static int ensure_staging(track2d_env *env, bool u8 = false) {
    const size_t E = (size_t)env->w.E, cells = (size_t)cells_of(env);
    int rc = T2D_OK;
    if (!env->d_actions) {
        rc = dev_alloc(env, &env->d_actions, 2 * E);
        if (rc == T2D_OK) rc = dev_alloc(env, &env->d_reward, 2 * E);
        if (rc == T2D_OK) rc = dev_alloc(env, &env->d_done, E);
        if (rc == T2D_OK) rc = dev_alloc(env, &env->d_mask, E);
    }
    if (rc == T2D_OK && !u8 && !env->d_obs) rc = dev_alloc(env, &env->d_obs, E * 2 * cells);
    if (rc == T2D_OK && u8 && !env->d_obs8) rc = dev_alloc(env, &env->d_obs8, E * 2 * cells + 16);
    return rc;
}

template <typename ObsT>
static int step_host_t(track2d_env *env, const int32_t *actions_host, ObsT *obs_host, float *reward_host, uint8_t *done_host) {
    constexpr bool u8 = sizeof(ObsT) == 1;
    int rc = ensure_staging(env, u8);
    if (rc != T2D_OK) return rc;
    cudaStream_t s = env->own_stream;
    const size_t E = (size_t)env->w.E, cells = (size_t)cells_of(env);
    ObsT *d_obs = u8 ? (ObsT *)env->d_obs8 : (ObsT *)env->d_obs;
    T2D_CUDA(cudaMemcpyAsync(env->d_actions, actions_host, 2 * E * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    rc = do_step<ObsT>(env, env->d_actions, d_obs, env->d_reward, env->d_done, s);
    if (rc != T2D_OK) return rc;
    if (obs_host) T2D_CUDA(cudaMemcpyAsync(obs_host, d_obs, E * 2 * cells * sizeof(ObsT), cudaMemcpyDeviceToHost, s));
    if (reward_host) T2D_CUDA(cudaMemcpyAsync(reward_host, env->d_reward, 2 * E * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (done_host) T2D_CUDA(cudaMemcpyAsync(done_host, env->d_done, E, cudaMemcpyDeviceToHost, s));
    T2D_CUDA(cudaStreamSynchronize(s));
    return T2D_OK;
}

// The same transition with the observation D2H issued in n_chunks pieces of consecutive envs; returns after ENQUEUEING.  Chunk c (and,
// with chunk 0, the reward / done arrays) is in the host buffers once track2d_host_chunk_wait(env, c) returns, so the consumer can
// process / re-upload chunk c while later chunks are still on the bus (PCIe is full duplex).
template <typename ObsT>
static int step_host_begin_t(track2d_env *env, const int32_t *actions_host, ObsT *obs_host, float *reward_host, uint8_t *done_host, int n_chunks) {
    constexpr bool u8 = sizeof(ObsT) == 1;
    int rc = ensure_staging(env, u8);
    if (rc != T2D_OK) return rc;
    while (env->n_chunk_ev < n_chunks) {
        T2D_CUDA(cudaEventCreateWithFlags(&env->chunk_ev[env->n_chunk_ev], cudaEventDisableTiming));
        env->n_chunk_ev++;
    }
    cudaStream_t s = env->own_stream;
    const size_t E = (size_t)env->w.E, cells = (size_t)cells_of(env), per_env = 2 * cells;
    ObsT *d_obs = u8 ? (ObsT *)env->d_obs8 : (ObsT *)env->d_obs;
    T2D_CUDA(cudaMemcpyAsync(env->d_actions, actions_host, 2 * E * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    rc = do_step<ObsT>(env, env->d_actions, d_obs, env->d_reward, env->d_done, s);
    if (rc != T2D_OK) return rc;
    T2D_CUDA(cudaMemcpyAsync(reward_host, env->d_reward, 2 * E * sizeof(float), cudaMemcpyDeviceToHost, s));
    T2D_CUDA(cudaMemcpyAsync(done_host, env->d_done, E, cudaMemcpyDeviceToHost, s));
    for (int c = 0; c < n_chunks; c++) {
        const size_t e0 = E * (size_t)c / (size_t)n_chunks, e1 = E * (size_t)(c + 1) / (size_t)n_chunks;
        if (e1 > e0)
            T2D_CUDA(cudaMemcpyAsync(obs_host + e0 * per_env, d_obs + e0 * per_env, (e1 - e0) * per_env * sizeof(ObsT), cudaMemcpyDeviceToHost, s));
        T2D_CUDA(cudaEventRecord(env->chunk_ev[c], s));
    }
    return T2D_OK;
}

template <typename ObsT>
static int reset_host_t(track2d_env *env, const uint8_t *mask_host, ObsT *obs_host) {
    constexpr bool u8 = sizeof(ObsT) == 1;
    int rc = ensure_staging(env, u8);
    if (rc != T2D_OK) return rc;
    cudaStream_t s = env->own_stream;
    const size_t E = (size_t)env->w.E, cells = (size_t)cells_of(env);
    ObsT *d_obs = u8 ? (ObsT *)env->d_obs8 : (ObsT *)env->d_obs;
    if (mask_host) T2D_CUDA(cudaMemcpyAsync(env->d_mask, mask_host, E, cudaMemcpyHostToDevice, s));
    rc = do_reset<ObsT>(env, mask_host ? env->d_mask : nullptr, d_obs, 0, s);
    if (rc != T2D_OK) return rc;
    if (obs_host) T2D_CUDA(cudaMemcpyAsync(obs_host, d_obs, E * 2 * cells * sizeof(ObsT), cudaMemcpyDeviceToHost, s));
    T2D_CUDA(cudaStreamSynchronize(s));
    return T2D_OK;
}

extern "C" {

int track2d_reset_host(track2d_env *env, const uint8_t *mask_host, float *obs_host) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return reset_host_t<float>(env, mask_host, obs_host);
}
int track2d_reset_host_u8(track2d_env *env, const uint8_t *mask_host, uint8_t *obs_host) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    return reset_host_t<uint8_t>(env, mask_host, obs_host);
}
int track2d_step_host(track2d_env *env, const int32_t *actions_host, float *obs_host, float *reward_host, uint8_t *done_host) {
    T2D_REQUIRE(env, "null handle");
    T2D_REQUIRE(actions_host, "step_host: actions required");
    DeviceGuard guard(env->cfg.device);
    return step_host_t<float>(env, actions_host, obs_host, reward_host, done_host);
}
int track2d_step_host_u8(track2d_env *env, const int32_t *actions_host, uint8_t *obs_host, float *reward_host, uint8_t *done_host) {
    T2D_REQUIRE(env, "null handle");
    T2D_REQUIRE(actions_host, "step_host_u8: actions required");
    DeviceGuard guard(env->cfg.device);
    return step_host_t<uint8_t>(env, actions_host, obs_host, reward_host, done_host);
}

int track2d_join(track2d_env *env, void *stream) {
    T2D_REQUIRE(env, "null handle");
    if (!env->standby || env->since_join == 0) return T2D_OK;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, env->ev_side[(env->since_join - 1) % 8], 0));
    env->since_join = 0;
    return T2D_OK;
}

int track2d_step_host_begin(track2d_env *env, const int32_t *actions_host, void *obs_host, int32_t obs_is_u8, float *reward_host, uint8_t *done_host,
                            int32_t n_chunks) {
    T2D_REQUIRE(env, "null handle");
    T2D_REQUIRE(actions_host && obs_host && reward_host && done_host, "step_host_begin: all four host buffers are required");
    T2D_REQUIRE(n_chunks >= 1 && n_chunks <= 16, "step_host_begin: 1..16 chunks");
    DeviceGuard guard(env->cfg.device);
    return obs_is_u8 ? step_host_begin_t<uint8_t>(env, actions_host, (uint8_t *)obs_host, reward_host, done_host, n_chunks)
                     : step_host_begin_t<float>(env, actions_host, (float *)obs_host, reward_host, done_host, n_chunks);
}
int track2d_host_chunk_wait(track2d_env *env, int32_t chunk) {
    T2D_REQUIRE(env, "null handle");
    T2D_REQUIRE(chunk >= 0 && chunk < env->n_chunk_ev, "host_chunk_wait: no such chunk (call track2d_step_host_begin first)");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaEventSynchronize(env->chunk_ev[chunk]));
    return T2D_OK;
}
// the same dependency without the host in the middle: work enqueued on `stream` after this call starts once chunk `chunk` is in the
// host buffers (e.g. its re-upload: the H2D engine then runs one chunk behind the D2H engine with no host wake-up between them)
int track2d_host_chunk_wait_stream(track2d_env *env, int32_t chunk, void *stream) {
    T2D_REQUIRE(env, "null handle");
    T2D_REQUIRE(chunk >= 0 && chunk < env->n_chunk_ev, "host_chunk_wait_stream: no such chunk (call track2d_step_host_begin first)");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, env->chunk_ev[chunk], 0));
    return T2D_OK;
}

// ---- state read-back / injection ---------------------------------------------------------------------
int track2d_get_maps(track2d_env *env, int32_t first, int32_t count, uint8_t *maze_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> words((size_t)count * T2D_MAP_WORDS);
    T2D_CUDA(cudaMemcpy(words.data(), env->w.maps + (size_t)first * T2D_MAP_WORDS, words.size() * 4, cudaMemcpyDeviceToHost));
    const int H = env->w.H, W = env->w.W;
    for (int i = 0; i < count; i++) unpack_map(words.data() + (size_t)i * T2D_MAP_WORDS, H, W, maze_host + (size_t)i * H * W);
    return T2D_OK;
}

int track2d_set_maps(track2d_env *env, int32_t first, int32_t count, const uint8_t *maze_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> words((size_t)count * T2D_MAP_WORDS);
    const int H = env->w.H, W = env->w.W;
    for (int i = 0; i < count; i++) pack_map(maze_host + (size_t)i * H * W, H, W, words.data() + (size_t)i * T2D_MAP_WORDS);
    T2D_CUDA(cudaMemcpy(env->w.maps + (size_t)first * T2D_MAP_WORDS, words.data(), words.size() * 4, cudaMemcpyHostToDevice));
    return T2D_OK;
}

int track2d_get_agents(track2d_env *env, int32_t first, int32_t count, int32_t *pos_host, int32_t *counters_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> p(count), c(count);
    T2D_CUDA(cudaMemcpy(p.data(), env->w.pos + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
    T2D_CUDA(cudaMemcpy(c.data(), env->w.ctr + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; i++) {
        if (pos_host) {
            pos_host[4 * i + 0] = p[i] & 255; pos_host[4 * i + 1] = (p[i] >> 8) & 255;
            pos_host[4 * i + 2] = (p[i] >> 16) & 255; pos_host[4 * i + 3] = p[i] >> 24;
        }
        if (counters_host) { counters_host[2 * i] = c[i] & 0xFFFF; counters_host[2 * i + 1] = c[i] >> 16; }
    }
    return T2D_OK;
}

int track2d_set_agents(track2d_env *env, int32_t first, int32_t count, const int32_t *pos_host, const int32_t *counters_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> v(count);
    if (pos_host) {
        for (int i = 0; i < count; i++) {
            for (int k = 0; k < 4; k++) {
                int lim = (k & 1) ? env->w.W : env->w.H;
                if (pos_host[4 * i + k] < 1 || pos_host[4 * i + k] > lim - 2) {
                    t2d_set_error("set_agents: position outside the map interior");
                    return T2D_E_INVALID;
                }
            }
            v[i] = (uint32_t)pos_host[4 * i] | ((uint32_t)pos_host[4 * i + 1] << 8) | ((uint32_t)pos_host[4 * i + 2] << 16) | ((uint32_t)pos_host[4 * i + 3] << 24);
        }
        T2D_CUDA(cudaMemcpy(env->w.pos + first, v.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    }
    if (counters_host) {
        for (int i = 0; i < count; i++) v[i] = ((uint32_t)counters_host[2 * i] & 0xFFFFu) | ((uint32_t)counters_host[2 * i + 1] << 16);
        T2D_CUDA(cudaMemcpy(env->w.ctr + first, v.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    }
    env->was_reset = true; // an injected state is steppable
    return T2D_OK;
}

int track2d_get_goals(track2d_env *env, int32_t first, int32_t count, int32_t *goals_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> g(count);
    T2D_CUDA(cudaMemcpy(g.data(), env->w.goals + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; i++)
        for (int k = 0; k < 4; k++) goals_host[4 * i + k] = (g[i] >> (8 * k)) & 255;
    return T2D_OK;
}

int track2d_get_ram(track2d_env *env, int32_t first, int32_t count, int32_t *plan_host, int32_t *len_host, int32_t *idx_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> r(count);
    T2D_CUDA(cudaMemcpy(r.data(), env->w.ram + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; i++) {
        for (int k = 0; k < TRACK2D_RAM_MAXPLAN; k++) plan_host[TRACK2D_RAM_MAXPLAN * i + k] = (r[i] >> (2 * k)) & 3;
        len_host[i] = (r[i] >> 18) & 15;
        idx_host[i] = (r[i] >> 22) & 15;
    }
    return T2D_OK;
}

int track2d_set_ram(track2d_env *env, int32_t first, int32_t count, const int32_t *plan_host, const int32_t *len_host, const int32_t *idx_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint32_t> r(count);
    for (int i = 0; i < count; i++) {
        T2D_REQUIRE(len_host[i] >= 1 && len_host[i] <= TRACK2D_RAM_MAXPLAN && idx_host[i] >= 0 && idx_host[i] < len_host[i], "set_ram: bad plan length/index");
        uint32_t bits = 0;
        for (int k = 0; k < len_host[i]; k++) bits |= ((uint32_t)plan_host[TRACK2D_RAM_MAXPLAN * i + k] & 3u) << (2 * k);
        r[i] = bits | ((uint32_t)len_host[i] << 18) | ((uint32_t)idx_host[i] << 22);
    }
    T2D_CUDA(cudaMemcpy(env->w.ram + first, r.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    return T2D_OK;
}

int track2d_get_nav(track2d_env *env, int32_t first, int32_t count, int32_t *plan_host, int32_t *len_host, int32_t *idx_host, int32_t *goal_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    T2D_REQUIRE(env->w.nav_plan, "get_nav: not a Nav/RPF env");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint8_t> plan((size_t)count * T2D_NAV_PLAN_BYTES);
    std::vector<uint32_t> meta(count), goal(count);
    T2D_CUDA(cudaMemcpy(plan.data(), env->w.nav_plan + (size_t)first * T2D_NAV_PLAN_BYTES, plan.size(), cudaMemcpyDeviceToHost));
    T2D_CUDA(cudaMemcpy(meta.data(), env->w.nav_meta + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
    T2D_CUDA(cudaMemcpy(goal.data(), env->w.nav_goal + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; i++) {
        // plan-ahead handles keep the plan as a ring whose length and cursor only grow: report it re-based at the cursor
        const int shift = env->w.async_nav ? (int)(meta[i] >> 16) : 0;
        if (plan_host)
            for (int k = 0; k < TRACK2D_NAV_MAXPLAN; k++) {
                const int q = (k + shift) & (TRACK2D_NAV_MAXPLAN - 1);
                plan_host[(size_t)TRACK2D_NAV_MAXPLAN * i + k] = (plan[(size_t)i * T2D_NAV_PLAN_BYTES + (q >> 2)] >> (2 * (q & 3))) & 3;
            }
        len_host[i] = (int)((meta[i] & 0xFFFF) - (uint32_t)shift) & 0xFFFF;
        idx_host[i] = (int)(meta[i] >> 16) - shift;
        if (goal_host) { goal_host[2 * i] = goal[i] & 255; goal_host[2 * i + 1] = (goal[i] >> 8) & 255; }
    }
    return T2D_OK;
}

int track2d_set_nav(track2d_env *env, int32_t first, int32_t count, const int32_t *plan_host, const int32_t *len_host, const int32_t *idx_host, const int32_t *goal_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    T2D_REQUIRE(env->w.nav_plan, "set_nav: not a Nav/RPF env");
    if (env->w.async_nav) {
        t2d_set_error("set_nav: this handle plans ahead of time (Philox + auto-reset); create it without T2D_FLAG_PLAN_AHEAD to inject plans");
        return T2D_E_UNSUPPORTED;
    }
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint8_t> plan((size_t)count * T2D_NAV_PLAN_BYTES, 0);
    std::vector<uint32_t> meta(count), goal(count);
    for (int i = 0; i < count; i++) {
        T2D_REQUIRE(len_host[i] >= 0 && len_host[i] <= TRACK2D_NAV_MAXPLAN && idx_host[i] >= 0, "set_nav: bad plan length/index");
        for (int k = 0; k < len_host[i]; k++)
            plan[(size_t)i * T2D_NAV_PLAN_BYTES + (k >> 2)] |= (uint8_t)((plan_host[(size_t)TRACK2D_NAV_MAXPLAN * i + k] & 3) << (2 * (k & 3)));
        meta[i] = (uint32_t)len_host[i] | ((uint32_t)idx_host[i] << 16);
        goal[i] = (uint32_t)goal_host[2 * i] | ((uint32_t)goal_host[2 * i + 1] << 8);
    }
    T2D_CUDA(cudaMemcpy(env->w.nav_plan + (size_t)first * T2D_NAV_PLAN_BYTES, plan.data(), plan.size(), cudaMemcpyHostToDevice));
    T2D_CUDA(cudaMemcpy(env->w.nav_meta + first, meta.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    T2D_CUDA(cudaMemcpy(env->w.nav_goal + first, goal.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    return T2D_OK;
}

int track2d_get_rewards_f64(track2d_env *env, int32_t first, int32_t count, double *rewards_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    T2D_REQUIRE(env->w.rew64, "get_rewards_f64: handle was created without T2D_FLAG_KEEP_F64");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    T2D_CUDA(cudaMemcpy(rewards_host, env->w.rew64 + 2 * (size_t)first, (size_t)count * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    return T2D_OK;
}

int track2d_astar_solve(track2d_env *env, int32_t first, int32_t count, const int32_t *start_host, const int32_t *goal_host, int32_t *plan_host,
                        int32_t *len_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    T2D_REQUIRE(env->w.nav_plan, "astar_solve: not a Nav/RPF env");
    T2D_REQUIRE(start_host && goal_host && len_host, "astar_solve: start, goal and len are required");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<int32_t> sg((size_t)count * 4);
    for (int i = 0; i < count; i++) {
        sg[4 * i] = start_host[2 * i]; sg[4 * i + 1] = start_host[2 * i + 1];
        sg[4 * i + 2] = goal_host[2 * i]; sg[4 * i + 3] = goal_host[2 * i + 1];
        for (int k = 0; k < 4; k++)
            T2D_REQUIRE(sg[4 * i + k] >= 0 && sg[4 * i + k] < ((k & 1) ? env->w.W : env->w.H), "astar_solve: start / goal outside the map");
    }
    int32_t *dbuf = nullptr;
    T2D_CUDA(cudaMalloc(&dbuf, (size_t)count * 5 * sizeof(int32_t)));
    cudaError_t err = cudaMemcpy(dbuf, sg.data(), sg.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        t2d_count_launches(1);
        err = t2d_launch_astar_direct(env->w, first, count, dbuf, dbuf + (size_t)count * 4, 0);
    }
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err == cudaSuccess) err = cudaMemcpy(len_host, dbuf + (size_t)count * 4, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaFree(dbuf);
    T2D_CUDA(err);
    if (plan_host) {
        std::vector<int32_t> ln(count), idx(count);
        return track2d_get_nav(env, first, count, plan_host, ln.data(), idx.data(), nullptr);
    }
    return T2D_OK;
}

int track2d_get_target_actions(track2d_env *env, int32_t first, int32_t count, int32_t *actions_host) {
    int rc = check_range(env, first, count);
    if (rc != T2D_OK) return rc;
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    std::vector<uint8_t> a(count);
    T2D_CUDA(cudaMemcpy(a.data(), env->w.tgt_act + first, (size_t)count, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; i++) actions_host[i] = a[i];
    return T2D_OK;
}

int track2d_seed_env(track2d_env *env, int32_t index, uint32_t seed) {
    int rc = check_range(env, index, 1);
    if (rc != T2D_OK) return rc;
    T2D_REQUIRE(env->w.mt_key, "seed_env: handle is not in T2D_RNG_NUMPY mode");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    T2D_CUDA(t2d_launch_seed_numpy(env->w, index, 1, seed, 0, 0));
    T2D_CUDA(cudaDeviceSynchronize());
    return T2D_OK;
}

int track2d_get_rng_numpy(track2d_env *env, int32_t index, uint32_t *key_host, int32_t *pos_host) {
    int rc = check_range(env, index, 1);
    if (rc != T2D_OK) return rc;
    T2D_REQUIRE(env->w.mt_key, "get_rng_numpy: handle is not in T2D_RNG_NUMPY mode");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    T2D_CUDA(cudaMemcpy(key_host, env->w.mt_key + (size_t)index * T2D_MT_N, T2D_MT_N * 4, cudaMemcpyDeviceToHost));
    T2D_CUDA(cudaMemcpy(pos_host, env->w.mt_pos + index, 4, cudaMemcpyDeviceToHost));
    return T2D_OK;
}

int track2d_get_status(track2d_env *env, uint32_t *status_host, void *stream) {
    T2D_REQUIRE(env && status_host, "get_status: null argument");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaMemcpyAsync(status_host, env->w.status, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    T2D_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return T2D_OK;
}

int track2d_get_counters(track2d_env *env, uint64_t *episodes_done, uint64_t *steps_done) {
    T2D_REQUIRE(env, "null handle");
    DeviceGuard guard(env->cfg.device);
    T2D_CUDA(cudaDeviceSynchronize());
    unsigned long long st[2];
    T2D_CUDA(cudaMemcpy(st, env->w.stats, sizeof(st), cudaMemcpyDeviceToHost));
    if (episodes_done) *episodes_done = st[0];
    if (steps_done) *steps_done = env->steps_done;
    return T2D_OK;
}

} // extern "C"
