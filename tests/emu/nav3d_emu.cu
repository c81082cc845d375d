// nav3d_emu.cu — DEBUGGING AID, tests only.  Compiles the device logic of csrc/nav3d_core.cuh for the HOST with one lane
// per env (G = 1) so that the packed-representation logic can be checked against the oracle in the GPU-less build
// container.  It is not linked into libnav3d_b200.so and nothing in the product imports it; it says nothing about the
// concurrency of the real kernels, which only the -m gpu tests exercise.
#include "../../3d-navigation-reinforcement-learning_b200/csrc/nav3d_core.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace nav3d;

struct Emu {
    EngineParams P{};
    std::vector<RoomDev> rooms;
    std::vector<uint16_t> occz;
    std::vector<unsigned long long> occ64;
    std::vector<uint32_t> free_cells;
    std::vector<EnvState> states;
    std::vector<uint8_t> know;
    float lut[nav3d::kLutSize];
    std::vector<float> dist_lut;
    bool simple = false;
};

extern "C" {

void *emu_create(int n_envs, int L, double crash, unsigned long long seed, unsigned env_id0, int auto_reset, int n_rooms,
                 const int *dims, const int8_t *dense, const unsigned *dense_off, int wall_code, int simple, double cell_size) {
    Emu *e = new Emu();
    e->simple = simple != 0;
    size_t max_s = 0, max_c = 0;
    for (int r = 0; r < n_rooms; r++) {
        const int W = dims[3 * r], D = dims[3 * r + 1], H = dims[3 * r + 2];
        const int8_t *g = dense + dense_off[r];
        RoomDev R{};
        R.W = W; R.D = D; R.H = H; R.ntx = (W + 3) / 4; R.nty = (D + 3) / 4; R.nbz = (H + 1) / 2;
        R.occz_off = (uint32_t)e->occz.size();
        for (int x = 0; x < W; x++) for (int y = 0; y < D; y++) {
            uint16_t b = 0;
            for (int z = 0; z < H; z++) if (g[(x * D + y) * H + z] == wall_code) b |= 1u << z;
            e->occz.push_back(b);
        }
        R.occx_off = (uint32_t)e->occ64.size();
        for (int y = 0; y < D; y++) for (int z = 0; z < H; z++) {
            unsigned long long b = 0;
            for (int x = 0; x < W; x++) if (g[(x * D + y) * H + z] == wall_code) b |= 1ull << x;
            e->occ64.push_back(b);
        }
        R.occy_off = (uint32_t)e->occ64.size();
        for (int x = 0; x < W; x++) for (int z = 0; z < H; z++) {
            unsigned long long b = 0;
            for (int y = 0; y < D; y++) if (g[(x * D + y) * H + z] == wall_code) b |= 1ull << y;
            e->occ64.push_back(b);
        }
        R.free_off = (uint32_t)e->free_cells.size();
        uint32_t nf = 0;
        for (int x = 1; x < W - 1; x++) for (int y = 1; y < D - 1; y++) for (int z = 1; z < H - 1; z++)
            if (g[(x * D + y) * H + z] != wall_code) { e->free_cells.push_back(x | (y << 8) | (z << 16)); nf++; }
        R.n_free = nf;
        e->rooms.push_back(R);
        max_s = std::max(max_s, (size_t)R.ntx * R.nty * 32);
        max_c = std::max(max_c, (size_t)R.ntx * R.nty * R.nbz * 32);
    }
    size_t c_off = (max_s + 127) / 128 * 128, stride = c_off + (max_c + 127) / 128 * 128;
    if (e->simple) { c_off = 0; stride = (2 * max_s + 127) / 128 * 128; }
    for (int c = 0; c <= L; c++) e->dist_lut.push_back((float)(std::nearbyint((double)c * cell_size * 100.0) / 100.0));
    e->states.assign((size_t)n_envs, EnvState{});
    e->know.assign(stride * (size_t)n_envs, 0xAB);     // poison: a reset must clear what it uses
    for (int i = 0; i < 23; i++) e->lut[i] = (float)i / 22.0f;
    for (int i = 0; i < 6; i++) e->lut[nav3d::kLutFifth + i] = (float)i / 5.0f;
    for (int i = 0; i < 32; i++) e->lut[nav3d::kLutDown + i] = (float)i / (float)L;
    EngineParams &P = e->P;
    P.rooms = e->rooms.data(); P.occz = e->occz.data(); P.occ64 = e->occ64.data(); P.free_cells = e->free_cells.data();
    P.states = e->states.data(); P.know = e->know.data(); P.env_stride = stride; P.c_off = (uint32_t)c_off;
    P.n_envs = n_envs; P.n_rooms = n_rooms; P.L = L; P.env_id0 = env_id0; P.seed_lo = (uint32_t)seed;
    P.seed_hi = (uint32_t)(seed >> 32); P.auto_reset = auto_reset; P.crash_penalty = crash;
    P.dist_lut = e->dist_lut.data(); P.obs_dim = e->simple ? 6 * L + 7 : kObsDim;
    return e;
}
void emu_destroy(void *h) { delete (Emu *)h; }
int emu_room_n_free(void *h, int r) { return (int)((Emu *)h)->rooms[r].n_free; }

void emu_reset(void *h, const int *env_ids, int n, const int *picks, float *obs) {
    Emu *e = (Emu *)h;
    for (int i = 0; i < n; i++) {
        const int env = env_ids ? env_ids[i] : i;
        const uint32_t ep = e->states[env].episode;
        float *orow = obs ? obs + (size_t)env * e->P.obs_dim : nullptr;
        if (e->simple) {
            if (picks) simple_reset_env<1>(e->P, env, 0, 0, (uint32_t)picks[3 * i], (uint32_t)picks[3 * i + 1], (uint32_t)picks[3 * i + 2], ep + 1, orow);
            else simple_reset_env_philox<1>(e->P, env, 0, 0, ep, orow);
            continue;
        }
        if (picks) reset_env<1>(e->P, env, 0, 0, (uint32_t)picks[2 * i], (uint32_t)picks[2 * i + 1], ep + 1, e->lut, orow);
        else reset_env_philox<1>(e->P, env, 0, 0, ep, e->lut, orow);
    }
}
void emu_step(void *h, const long long *actions, float *obs, float *reward, double *reward64, uint8_t *term,
              uint8_t *trunc, float *terminal_obs, void *episodes) {
    Emu *e = (Emu *)h;
    StepIO io;
    io.actions = actions; io.obs = obs; io.reward = reward; io.reward64 = reward64; io.terminated = term;
    io.truncated = trunc; io.terminal_obs = terminal_obs; io.episodes = episodes; io.env0 = 0; io.env_n = e->P.n_envs;
    if (e->simple) {
        for (int env = 0; env < e->P.n_envs; env++) simple_step_env<1>(e->P, io, env, 0, 0, (int)actions[env], env);
        return;
    }
    for (int env = 0; env < e->P.n_envs; env++) step_env<1, true>(e->P, io, env, 0, 0, (int)actions[env], e->lut, env);
}
void emu_get_state(void *h, int *out) {
    Emu *e = (Emu *)h;
    for (int env = 0; env < e->P.n_envs; env++) {
        const EnvState s = e->states[env];
        int *o = out + (size_t)env * 16;
        o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.facing; o[4] = s.visited_count; o[5] = s.bump_count;
        o[6] = s.step_count; o[7] = (s.flags & kNearWall) != 0; o[8] = (s.flags & kWasNearWall) != 0;
        o[9] = (s.flags & kLastBump) != 0; o[10] = (s.flags & kDone) != 0; o[11] = s.down; o[12] = s.last_action;
        o[13] = s.room; o[14] = s.episode; o[15] = s.ret_centi;
        if (e->simple) { o[7] = s.down; o[8] = s.pad0; o[9] = s.pad1; o[11] = 0; }
    }
}
void emu_get_grid(void *h, int env, int16_t *out) {
    Emu *e = (Emu *)h;
    const EnvState s = e->states[env];
    const RoomDev R = e->rooms[s.room];
    const uint8_t *envk = e->know.data() + (size_t)env * e->P.env_stride;
    const uint16_t *S = (const uint16_t *)envk;
    const uint8_t *C = envk + e->P.c_off;
    if (e->simple) {
        const uint32_t *K = (const uint32_t *)envk;
        for (int x = 0; x < R.W; x++) for (int y = 0; y < R.D; y++) for (int z = 0; z < R.H; z++)
            out[(x * R.D + y) * R.H + z] = (int16_t)(k2_code(K[s_index(R, x, y)], z) - 1);
        return;
    }
    for (int x = 0; x < R.W; x++) for (int y = 0; y < R.D; y++) for (int z = 0; z < R.H; z++) {
        const uint32_t sw = S[s_index(R, x, y)], ow = e->occz[R.occz_off + x * R.D + y];
        int16_t v = -1;
        if ((sw >> z) & 1u) v = ((ow >> z) & 1u) ? -2 : (int16_t)C[c_index(R, x, y, z)];
        out[(x * R.D + y) * R.H + z] = v;
    }
}
}
