// nav3d_core.cuh — per-environment logic of the batched CubicEnv step, written once and instantiated for a group of
// G cooperating lanes (G = 1 .. 32).  Everything here is device code for sm_100a; the functions are also marked
// __host__ so that tests/emu can run the SAME source with G = 1 on the CPU as a debugging aid (never shipped,
// never on the product path).
//
// Reference being replaced (semantics, not structure): envs/CubicEnv.py of Noimps/3D-Navigation-Reinforcement-Learning
//   step :110-132, do_action :134-166, compute_reward :169-224, _get_3d_local_map :229-251, get_obs :254-312,
//   _mark_visited :322-343, _sense_direction :345-397, reset :77-108, load_room's start pick :450-462.
//
// Data model.  The reference keeps one int64 per voxel per env (internal_grid: -2 known wall, -1 unknown,
// 0 seen, >=1 visit counter).  Here that grid is factored into
//   * the static room occupancy, shared by all envs, bit-packed in three orientations so that each of the six
//     axis-aligned rays is ONE 64-bit word + a bit scan (no per-cell march):
//        occz[x][y] : u16, bit z        occx[y][z] : u64, bit x        occy[x][z] : u64, bit y
//   * a per-env "seen" bit volume S: u16 per (x,y) column (bit z), stored in 4x4-column tiles of 32 B (one sector)
//   * a per-env visit-count volume C: u8 per cell, saturating at 255, stored in 4x4x2 bricks of 32 B
//   internal_grid[c] == (!S[c] ? -1 : occ[c] ? -2 : C[c]).  Observations clip counters at 20 and the reward at 25
//   (CubicEnv.py:273-274, :180), so saturation at 255 changes no output.
// No lane ever depends on another lane's memory writes inside a step (the window gather re-derives the freshly seen
// bits from the ray extents in registers), so a step needs no intra-group synchronisation; only a reset does.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define NAV3D_HD __host__ __device__ __forceinline__

namespace nav3d {

constexpr int kObsDim = 80;
// Shared look-up table of the step kernels: [0, 23) (m + 2) / 22 for the window (CubicEnv.py:273-275); [24, 30) k / 5 for
// last_action (:284); [32, 64) d / L for cells_insight_down when L <= 31 (:287).  All entries are correctly rounded f32
// quotients, i.e. the same bits as computing them in place.
constexpr int kLutSize = 64, kLutFifth = 24, kLutDown = 32;
constexpr uint32_t kStreamReset = 0x52455345u;   // include/nav3d.h "Random streams"
constexpr uint32_t kStreamAction = 0x41435449u;

// flag bits of EnvState::flags
constexpr uint32_t kNearWall = 1u, kWasNearWall = 2u, kLastBump = 4u, kDone = 8u;

struct alignas(32) RoomDev {       // one per room, read-only after nav3d_load_rooms
    uint16_t W, D, H, ntx;         // dims; ntx = ceil(W/4) tiles along x
    uint16_t nty, nbz;             // nty = ceil(D/4); nbz = ceil(H/2) bricks along z
    uint32_t n_free;               // total_free_cells == max_steps (CubicEnv.py:450-459)
    uint32_t occz_off;             // u16 index into occz pool, [x][y]
    uint32_t occx_off;             // u64 index into occ64 pool, [y][z], bit x
    uint32_t occy_off;             // u64 index into occ64 pool, [x][z], bit y
    uint32_t free_off;             // u32 index into the free-cell pool (x | y<<8 | z<<16, reference scan order)
};
static_assert(sizeof(RoomDev) == 32, "RoomDev must be one 32-byte sector");

struct alignas(32) EnvState {      // one 32-byte sector per env: read once, written once per step
    uint8_t x, y, z, facing;
    uint8_t last_action, flags, down, pad0;
    uint32_t step_count, visited_count, bump_count;
    int32_t ret_centi;             // running episode return in 1/100 units, crash penalties excluded
    uint32_t episode;              // number of resets so far == index into the Philox reset stream
    uint16_t room, pad1;
};
static_assert(sizeof(EnvState) == 32, "EnvState must be one 32-byte sector");

struct EngineParams {
    const RoomDev *rooms;
    const uint16_t *occz;
    const unsigned long long *occ64;
    const uint32_t *free_cells;
    EnvState *states;
    uint8_t *know;                 // per-env knowledge storage: [S tiles | C bricks]
    unsigned long long env_stride; // bytes per env in `know`
    uint32_t c_off;                // byte offset of the C bricks inside an env block
    int32_t n_envs, n_rooms, L;
    uint32_t env_id0, seed_lo, seed_hi;
    int32_t auto_reset;
    double crash_penalty;
    const float *dist_lut;         // simpleEnv: round(count * cell_size, 2) as f32, count = 0..L (simpleEnv.py:337)
    int32_t obs_dim;               // 80 (CubicEnv) or 6L+7 (simpleEnv)
};

struct StepIO {                    // per-launch output pointers (any may be null except obs)
    const long long *actions;
    float *obs;
    float *reward;
    double *reward64;
    uint8_t *terminated, *truncated;
    float *terminal_obs;
    void *episodes;                // nav3d_episode*
    int env0, env_n;               // the launch covers envs [env0, env0 + env_n); outputs are indexed by the env itself
};

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
NAV3D_HD int ffs64(unsigned long long v) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v);
#else
    return v ? __builtin_ctzll(v) + 1 : 0;
#endif
}
NAV3D_HD int clz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
NAV3D_HD float fdiv_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
template <typename T> NAV3D_HD T ldg(const T *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
NAV3D_HD int imin(int a, int b) { return a < b ? a : b; }
// Streaming (evict-first) stores for data that is written once and not read again by this engine (observations, rewards,
// flags): keeps them from evicting the per-env knowledge lines that the NEXT step will touch again from L2.
NAV3D_HD void store_stream(float4 *p, float4 v) {
#ifdef __CUDA_ARCH__
    __stcs(p, v);
#else
    *p = v;
#endif
}
NAV3D_HD void store_stream(float *p, float v) {
#ifdef __CUDA_ARCH__
    __stcs(p, v);
#else
    *p = v;
#endif
}

template <int G> NAV3D_HD void group_sync(int lane_in_warp) {
#ifdef __CUDA_ARCH__
    if (G == 32) __syncwarp();
    else if (G > 1) {
        unsigned m = (G >= 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane_in_warp & ~(G - 1));
        __syncwarp(m);
    }
#else
    (void)lane_in_warp;
#endif
}

// Philox4x32-10 (Salmon et al., SC'11)
NAV3D_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t &o0, uint32_t &o1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1;
}
NAV3D_HD void philox4x32_10_4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *o) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
NAV3D_HD uint32_t mulhi_range(uint32_t u, uint32_t n) { return (uint32_t)(((unsigned long long)u * n) >> 32); }

// ---------------------------------------------------------------------------------------------------------------
// knowledge-storage addressing
// ---------------------------------------------------------------------------------------------------------------
NAV3D_HD uint32_t s_index(const RoomDev &R, int x, int y) {          // u16 index of column (x,y) in the S tiles
    return (uint32_t)((((y >> 2) * R.ntx + (x >> 2)) << 4) + ((x & 3) << 2) + (y & 3));
}
NAV3D_HD uint32_t c_index(const RoomDev &R, int x, int y, int z) {   // byte index of cell (x,y,z) in the C bricks
    return (uint32_t)(((((y >> 2) * R.ntx + (x >> 2)) * R.nbz + (z >> 1)) << 5) + ((x & 3) << 3) + ((y & 3) << 1) + (z & 1));
}
NAV3D_HD uint32_t s_bytes(const RoomDev &R) { return (uint32_t)R.ntx * R.nty * 32u; }
NAV3D_HD uint32_t c_bytes(const RoomDev &R) { return (uint32_t)R.ntx * R.nty * R.nbz * 32u; }

// ---------------------------------------------------------------------------------------------------------------
// rays: _sense_direction (CubicEnv.py:345-397) for all six directions at once, from the packed occupancy
// ---------------------------------------------------------------------------------------------------------------
struct Rays {
    int x0, x1, y0, y1;        // inclusive extents of the cells whose knowledge the +-x / +-y rays touch
    uint32_t zmask;            // bits of the centre column touched by the +-z rays (centre bit included)
    int near_wall;             // a wall at distance 1 in any direction (:378-379)
    int down;                  // free cells seen below (:394-395)
    uint32_t blocked6;         // bit d set: a move in direction d bumps (wall or room edge at distance 1);
                               // d = 0 +x, 1 -x, 2 +y, 3 -y, 4 +z, 5 -z.  Cached in the record for the next step's move.
};

NAV3D_HD int ffs32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __ffs((int)v);
#else
    return v ? __builtin_ctz(v) + 1 : 0;
#endif
}
NAV3D_HD int clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
// cells p+1 .. p+n along increasing bit index of w.  SHORT (ray length <= 31): only the 32 cells next to p can matter, so
// the scan runs on one 32-bit word (a wall further away reads as "none", which is what f > n means anyway).
template <bool SHORT>
NAV3D_HD void ray_up(unsigned long long w, int p, int n, int &ext, int &nfree, int &near) {
    ext = 0; nfree = 0; near = 0;
    if (n <= 0) return;
    unsigned long long m = w >> (p + 1);                 // n > 0 implies p + 1 <= 63
    int f = SHORT ? ffs32((uint32_t)m) : ffs64(m);       // 1-based distance of the first wall, 0 = none
    if (f != 0 && f <= n) { ext = f; nfree = f - 1; near = (f == 1); }
    else { ext = n; nfree = n; }
}
// cells p-1 .. p-n along decreasing bit index of w
template <bool SHORT>
NAV3D_HD void ray_down(unsigned long long w, int p, int n, int &ext, int &nfree, int &near) {
    ext = 0; nfree = 0; near = 0;
    if (n <= 0) return;                                  // n > 0 implies 1 <= p <= 63
    unsigned long long m = w << (64 - p);                // bit 63 = cell p-1
    int f;
    if (SHORT) { const uint32_t hi = (uint32_t)(m >> 32); f = hi ? clz32(hi) + 1 : 0; }
    else f = m ? clz64(m) + 1 : 0;
    if (f != 0 && f <= n) { ext = f; nfree = f - 1; near = (f == 1); }
    else { ext = n; nfree = n; }
}

template <bool SHORT>
NAV3D_HD Rays cast_rays_t(const EngineParams &P, const RoomDev &R, int x, int y, int z) {
    const int L = P.L, W = R.W, D = R.D, H = R.H;
    unsigned long long wx = ldg(P.occ64 + R.occx_off + (uint32_t)(y * H + z));
    unsigned long long wy = ldg(P.occ64 + R.occy_off + (uint32_t)(x * H + z));
    unsigned long long wz = ldg(P.occz + R.occz_off + (uint32_t)(x * D + y));
    Rays r;
    int ext, nfree, near, any = 0, n;
    uint32_t blk = 0;
    n = imin(L, W - 1 - x); ray_up<SHORT>(wx, x, n, ext, nfree, near);   r.x1 = x + ext; any |= near; blk |= (uint32_t)(near | (n <= 0)) << 0;
    n = imin(L, x);         ray_down<SHORT>(wx, x, n, ext, nfree, near); r.x0 = x - ext; any |= near; blk |= (uint32_t)(near | (n <= 0)) << 1;
    n = imin(L, D - 1 - y); ray_up<SHORT>(wy, y, n, ext, nfree, near);   r.y1 = y + ext; any |= near; blk |= (uint32_t)(near | (n <= 0)) << 2;
    n = imin(L, y);         ray_down<SHORT>(wy, y, n, ext, nfree, near); r.y0 = y - ext; any |= near; blk |= (uint32_t)(near | (n <= 0)) << 3;
    int zu, zd;
    n = imin(L, H - 1 - z); ray_up<SHORT>(wz, z, n, zu, nfree, near);    any |= near; blk |= (uint32_t)(near | (n <= 0)) << 4;
    n = imin(L, z);         ray_down<SHORT>(wz, z, n, zd, nfree, near);  any |= near; blk |= (uint32_t)(near | (n <= 0)) << 5;
    r.down = nfree;
    r.zmask = ((2u << (z + zu)) - 1u) & ~((1u << (z - zd)) - 1u);
    r.near_wall = any;
    r.blocked6 = blk;
    return r;
}
NAV3D_HD Rays cast_rays(const EngineParams &P, const RoomDev &R, int x, int y, int z) {
    return P.L <= 31 ? cast_rays_t<true>(P, R, x, y, z) : cast_rays_t<false>(P, R, x, y, z);
}

// ---------------------------------------------------------------------------------------------------------------
// get_obs (CubicEnv.py:254-312): fold the rays into S, gather the 4x4x4 window, write the 80 floats
// ---------------------------------------------------------------------------------------------------------------
struct ObsScalars {
    int facing, last_action, was_near_wall, last_bump, down;
    uint32_t visited, total_free;
};

// 4 bits -> 4 byte masks (bit k -> byte k = 0xFF)
NAV3D_HD uint32_t expand4(uint32_t b) {
    const uint32_t m = (b * 0x00204081u) & 0x01010101u;      // the four partial products land on distinct bits: no carries
    return (m << 8) - m;
}
NAV3D_HD uint32_t vminu4_20(uint32_t v) {                    // per-byte min(v, 20)
#ifdef __CUDA_ARCH__
    return __vminu4(v, 0x14141414u);
#else
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) { uint32_t b = (v >> (8 * k)) & 0xffu; r |= (b < 20u ? b : 20u) << (8 * k); }
    return r;
#endif
}

// Lane -> window columns.  Column j = 4*dxi + dyi of the 4x4 (x,y) window, dyi = (lane & 3) + b*G (b < NY),
// dxi = (lane >> 2) + a*XS (a < NX); a batch is AB a-iterations (about four columns whose loads are in flight together).
template <int G> struct WinMap {
    static constexpr int NY = G >= 4 ? 1 : 4 / G;
    static constexpr int NX = G >= 16 ? 1 : (16 / G) / NY;
    static constexpr int XS = G >= 4 ? G / 4 : 1;
    static constexpr int AB = (4 / NY) < NX ? (4 / NY) : NX;
};
template <int G> struct WinBatch {                 // what one batch of window loads leaves in registers
    uint32_t sw[WinMap<G>::AB][WinMap<G>::NY];     // S word of the column
    uint32_t ow[WinMap<G>::AB][WinMap<G>::NY];     // occupancy word of the column
    unsigned long long cw[WinMap<G>::AB][WinMap<G>::NY];   // counters of the column for z-bricks zb0 .. zb0+2
    bool inb[WinMap<G>::AB][WinMap<G>::NY];
};

// Issue the loads of one batch (Step 2 of get_obs, :270).  Every index splits into an x half and a y half (see s_index /
// c_index), computed once per a / per b.
template <int G>
NAV3D_HD void window_load(const EngineParams &P, const RoomDev &R, const uint8_t *envk, int lane, int x, int y, int z,
                          int a0, WinBatch<G> &wb) {
    using M = WinMap<G>;
    const uint16_t *__restrict__ S = reinterpret_cast<const uint16_t *>(envk);
    const uint8_t *__restrict__ C = envk + P.c_off;
    const int nbz32 = R.nbz * 32, ntx16 = R.ntx * 16, cys = R.ntx * nbz32;
    const int zb0 = (z - 2) >> 1;                        // first z-brick of the window (may be -1)
    const bool odd = ((z - 2) & 1) != 0;
    const bool b0 = zb0 >= 0, b1 = zb0 + 1 < R.nbz, b2 = odd && zb0 + 2 < R.nbz;
    const int zoff = zb0 * 32;
#pragma unroll
    for (int a = 0; a < M::AB; a++) {
        const int dxi = (lane >> 2) + (a0 + a) * M::XS;
        const int cx = x + dxi - 2;
        const bool xin = dxi < 4 && cx >= 0 && cx < R.W;
        const int xs = ((cx >> 2) << 4) + ((cx & 3) << 2);
        const int xc = (cx >> 2) * nbz32 + ((cx & 3) << 3) + zoff;
        const int xo = cx * R.D;
#pragma unroll
        for (int b = 0; b < M::NY; b++) {
            const int cy = y + (lane & 3) + b * G - 2;
            const bool in = xin && cy >= 0 && cy < R.D;
            wb.inb[a][b] = in;
            wb.sw[a][b] = 0; wb.ow[a][b] = 0; wb.cw[a][b] = 0;
            if (in) {
                wb.sw[a][b] = S[xs + (cy >> 2) * ntx16 + (cy & 3)];
                wb.ow[a][b] = ldg(P.occz + R.occz_off + (uint32_t)(xo + cy));
                const uint16_t *cp = reinterpret_cast<const uint16_t *>(C + (xc + (cy >> 2) * cys + ((cy & 3) << 1)));
                unsigned long long w = 0;                              // bricks are 32 B = 16 u16 apart
                if (b0) w = cp[0];
                if (b1) w |= (unsigned long long)cp[16] << 16;
                if (b2) w |= (unsigned long long)cp[32] << 32;
                wb.cw[a][b] = w;
            }
        }
    }
}

// Turn one batch into observation floats: clip to [-2, 20], (m + 2) / 22 (:273-275), one float4 per column.
template <int G>
NAV3D_HD void window_store(const RoomDev &R, int lane, int x, int y, int z, int a0, const WinBatch<G> &wb, const Rays &r,
                           int centre_count, const float *lut, float *__restrict__ obs_row) {
    using M = WinMap<G>;
    const uint32_t zbit = 1u << z;
    const int zsh = ((z - 2) & 1) * 8;                   // bit offset of cell z-2 inside the column word
    const uint32_t zvalid = ((((1u << R.H) - 1u) << 2) >> z) & 15u;    // window cells inside [0, H)
    const float unknown = lut[1];
#pragma unroll
    for (int a = 0; a < M::AB; a++) {
        const int dxi = (lane >> 2) + (a0 + a) * M::XS;
        if (dxi >= 4) continue;
        const int cx = x + dxi - 2;
        const bool xray = cx >= r.x0 && cx <= r.x1;
#pragma unroll
        for (int b = 0; b < M::NY; b++) {
            const int dyi = (lane & 3) + b * G, cy = y + dyi - 2;
            float4 v = make_float4(unknown, unknown, unknown, unknown);
            if (wb.inb[a][b]) {
                uint32_t sbits = wb.sw[a][b];
                const bool centre_col = (dxi == 2 && dyi == 2);
                // the cells this step's rays see (they may not be in memory yet: marking happens after the gather)
                if (dyi == 2 && xray) sbits |= centre_col ? r.zmask : zbit;
                if (dxi == 2 && cy >= r.y0 && cy <= r.y1) sbits |= zbit;
                uint32_t c4 = (uint32_t)(wb.cw[a][b] >> zsh);                    // byte k = counter of cell z-2+k
                if (centre_col) c4 = (c4 & 0xff00ffffu) | ((uint32_t)centre_count << 16);
                c4 = vminu4_20(c4) + 0x02020202u;                               // clip at 20, +2 = LUT index of a free cell
                const uint32_t s4 = ((sbits << 2) >> z) & zvalid;               // seen, in range
                const uint32_t w4 = ((wb.ow[a][b] << 2) >> z) & s4;             // ... and a wall
                const uint32_t seen = expand4(s4), wall = expand4(w4);
                const uint32_t idx = (c4 & seen & ~wall) | (0x01010101u & ~seen);   // unknown -> 1, known wall -> 0
                v.x = lut[idx & 0xffu]; v.y = lut[(idx >> 8) & 0xffu];
                v.z = lut[(idx >> 16) & 0xffu]; v.w = lut[idx >> 24];
            }
            store_stream(reinterpret_cast<float4 *>(obs_row) + (dxi * 4 + dyi), v);
        }
    }
}

// Steps 3-6 of get_obs: the 9 scalars + zero padding = 4 more float4 (:279-307)
template <int G>
NAV3D_HD void write_scalars(const EngineParams &P, int lane, const ObsScalars &sc, const float *lut,
                            float *__restrict__ obs_row) {
    for (int j = 16 + lane; j < 20; j += G) {
        float4 v;
        if (j == 16) {
            v.x = sc.facing == 0 ? 1.f : 0.f; v.y = sc.facing == 1 ? 1.f : 0.f;
            v.z = sc.facing == 2 ? 1.f : 0.f; v.w = sc.facing == 3 ? 1.f : 0.f;
        } else if (j == 17) {
            // float(k)/5, count/L and visited/total are f64 quotients rounded to f32 in the reference (:284-291); for
            // integers below 2^24 that equals the correctly rounded f32 quotient (53 >= 2*24+2, Figueroa 1995).
            // (the two small quotients come from the shared table: same correctly rounded bits, no division sequence)
            v.x = lut[kLutFifth + sc.last_action];
            v.y = (float)sc.was_near_wall;
            v.z = (float)sc.last_bump;
            v.w = P.L <= 31 ? lut[kLutDown + sc.down] : fdiv_rn((float)sc.down, (float)P.L);
        } else if (j == 18) {
            v.x = fdiv_rn((float)sc.visited, (float)sc.total_free);
            v.y = v.z = v.w = 0.f;
        } else {
            v.x = v.y = v.z = v.w = 0.f;
        }
        store_stream(reinterpret_cast<float4 *>(obs_row) + j, v);
    }
}

// Step 1 of get_obs (:264-266): mark every cell the six rays examined as seen.  Runs after the gather in program order
// (the gather re-derives these bits from the ray extents) so that its stores do not fence the window loads.  The x run
// (centre column included, which also takes the +-z cells) and the y run are handled as one index space.
//
// `dir` = the move that brought the agent here (0 +x, 1 -x, 2 +y, 3 -y, 4 +z, 5 -z; -1 = none: a reset).  Every cell the
// agent has stood on had its rays marked when it was first visited, and the cell it just left lies one step back on the
// move's axis: that cell's rays along the axis already covered everything this cell's rays cover except the single cell at
// distance exactly L ahead.  So after a move along x (y) the x (y) run shrinks to that one far cell (if the ray reaches
// it) — the scattered tiles of a whole run become one.
template <int G>
NAV3D_HD void mark_seen(const RoomDev &R, uint8_t *envk, int lane, int x, int y, int z, const Rays &r, int dir, int L) {
    uint16_t *__restrict__ S = reinterpret_cast<uint16_t *>(envk);
    const uint32_t zbit = 1u << z;
    const int ntx16 = R.ntx * 16;
    int x0 = r.x0, x1 = r.x1, y0 = r.y0, y1 = r.y1, far = -1;
    if (dir == 0) { if (x1 - x == L) far = s_index(R, x1, y); x0 = x1 = x; }
    else if (dir == 1) { if (x - x0 == L) far = s_index(R, x0, y); x0 = x1 = x; }
    else if (dir == 2) { if (y1 - y == L) far = s_index(R, x, y1); y0 = y1 = y; }
    else if (dir == 3) { if (y - y0 == L) far = s_index(R, x, y0); y0 = y1 = y; }
    const int nx = x1 - x0 + 1, total = nx + (y1 - y0 + 1);
    const int ypart = (y >> 2) * ntx16 + (y & 3), xpart = ((x >> 2) << 4) + ((x & 3) << 2);
    constexpr int RC = G >= 16 ? 2 : 4;                 // cells per lane per chunk: loads of a chunk overlap
    uint32_t far_old = 0;
    const bool do_far = far >= 0 && lane == G - 1;
    if (do_far) far_old = S[far];
    for (int base = 0; base < total; base += G * RC) {
        int idx[RC];
        uint32_t old[RC], msk[RC];
#pragma unroll
        for (int q = 0; q < RC; q++) {
            const int i = base + q * G + lane;
            int id = -1;
            msk[q] = zbit;
            if (i < nx) {
                const int cx = x0 + i;
                id = ((cx >> 2) << 4) + ((cx & 3) << 2) + ypart;
                if (cx == x) msk[q] = r.zmask;
            } else if (i < total) {
                const int cy = y0 + (i - nx);
                if (cy != y) id = (cy >> 2) * ntx16 + (cy & 3) + xpart;       // centre column belongs to the x run
            }
            idx[q] = id;
            old[q] = id >= 0 ? (uint32_t)S[id] : 0u;
        }
#pragma unroll
        for (int q = 0; q < RC; q++) {
            const uint32_t n = old[q] | msk[q];
            if (idx[q] >= 0 && n != old[q]) S[idx[q]] = (uint16_t)n;
        }
    }
    if (do_far && (far_old | zbit) != far_old) S[far] = (uint16_t)(far_old | zbit);
}

// get_obs (CubicEnv.py:254-312) in one piece, for callers that have nothing to overlap with the window loads (reset).
template <int G>
NAV3D_HD void observe(const EngineParams &P, const RoomDev &R, uint8_t *envk, int lane, int x, int y, int z,
                      const Rays &r, int centre_count, bool write_seen, const ObsScalars &sc, const float *lut,
                      float *__restrict__ obs_row) {
    if (obs_row != nullptr) {
#pragma unroll
        for (int a0 = 0; a0 < WinMap<G>::NX; a0 += WinMap<G>::AB) {
            WinBatch<G> wb;
            window_load<G>(P, R, envk, lane, x, y, z, a0, wb);
            window_store<G>(R, lane, x, y, z, a0, wb, r, centre_count, lut, obs_row);
        }
        write_scalars<G>(P, lane, sc, lut, obs_row);
    }
    if (write_seen) mark_seen<G>(R, envk, lane, x, y, z, r, -1, P.L);
}

// ---------------------------------------------------------------------------------------------------------------
// reset (CubicEnv.py:77-108) of one env, room/start already chosen
// ---------------------------------------------------------------------------------------------------------------
template <int G>
NAV3D_HD void reset_env(const EngineParams &P, int env, int lane, int lane_in_warp, uint32_t room_idx, uint32_t k,
                        uint32_t episode_after, const float *lut, float *obs_row) {
    const RoomDev R = P.rooms[room_idx];
    uint8_t *envk = P.know + (unsigned long long)env * P.env_stride;
    // internal_grid = full(-1) (:84): nothing seen, nothing counted
    {
        uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        uint4 *s4 = reinterpret_cast<uint4 *>(envk);
        const uint32_t ns = s_bytes(R) >> 4;
        for (uint32_t i = lane; i < ns; i += G) s4[i] = zero;
        uint4 *c4 = reinterpret_cast<uint4 *>(envk + P.c_off);
        const uint32_t nc = c_bytes(R) >> 4;
        for (uint32_t i = lane; i < nc; i += G) c4[i] = zero;
    }
    group_sync<G>(lane_in_warp);
    const uint32_t cell = ldg(P.free_cells + R.free_off + k);
    const int x = cell & 0xff, y = (cell >> 8) & 0xff, z = (cell >> 16) & 0xff;
    if (lane == 0) envk[P.c_off + c_index(R, x, y, z)] = 1;           // internal_grid[start] = 1 (:85)
    const Rays r = cast_rays(P, R, x, y, z);                          // get_obs() inside reset (:108)
    ObsScalars sc;
    sc.facing = 0; sc.last_action = 0; sc.was_near_wall = 0; sc.last_bump = 0; sc.down = r.down;
    sc.visited = 1; sc.total_free = R.n_free;
    observe<G>(P, R, envk, lane, x, y, z, r, 1, true, sc, lut, obs_row);
    if (lane == 0) {
        EnvState st;
        st.x = (uint8_t)x; st.y = (uint8_t)y; st.z = (uint8_t)z; st.facing = 0;
        st.last_action = 0; st.flags = (uint8_t)(r.near_wall ? kNearWall : 0u); st.down = (uint8_t)r.down;
        st.pad0 = (uint8_t)r.blocked6;
        st.step_count = 0; st.visited_count = 1; st.bump_count = 0; st.ret_centi = 0;
        st.episode = episode_after; st.room = (uint16_t)room_idx; st.pad1 = 0;
        P.states[env] = st;
    }
}

template <int G>
NAV3D_HD void reset_env_philox(const EngineParams &P, int env, int lane, int lane_in_warp, uint32_t episode,
                               const float *lut, float *obs_row) {
    uint32_t u0, u1;
    philox4x32_10(P.env_id0 + (uint32_t)env, episode, 0u, kStreamReset, P.seed_lo, P.seed_hi, u0, u1);
    const uint32_t room = mulhi_range(u0, (uint32_t)P.n_rooms);       // random.choice(self.rooms) (:407)
    const uint32_t nf = ldg(&P.rooms[room].n_free);
    const uint32_t k = mulhi_range(u1, nf);                           // random.choice(possible_start_pose) (:462)
    reset_env<G>(P, env, lane, lane_in_warp, room, k, episode + 1u, lut, obs_row);
}

// ---------------------------------------------------------------------------------------------------------------
// step (CubicEnv.py:110-132) of one env
// ---------------------------------------------------------------------------------------------------------------
struct EpisodeRec { float episode_return; int32_t length, bumps, visited, total_free, room, terminated, truncated; };

// Returns true when the env finished its episode and must be reset by the caller (auto_reset and !INLINE_RESET).
// REG_STATE (fused multi-step rollouts): the env's record lives in the caller's registers (`rs`, every lane holds a
// copy and every lane computes the new one) instead of making a round trip through global memory each step; the
// terminated / truncated bits come back in `done_bits` and the flag arrays of `io` may be null.
template <int G, bool INLINE_RESET, bool REG_STATE = false>
NAV3D_HD bool step_env(const EngineParams &P, const StepIO &io, int env, int lane, int lane_in_warp, int action,
                       const float *lut, long long row /* row index for the output arrays */, EnvState *rs = nullptr,
                       uint32_t *done_bits = nullptr) {
    const EnvState st = REG_STATE ? *rs : P.states[env];
    const RoomDev R = P.rooms[st.room];
    uint8_t *envk = P.know + (unsigned long long)env * P.env_stride;
    uint8_t *C = envk + P.c_off;

    int a = action < 0 ? 0 : (action > 5 ? 5 : action);
    uint32_t flags = st.flags;
    if (flags & kNearWall) flags = (flags | kWasNearWall) & ~kNearWall;          // :111-113
    const uint32_t step_count = st.step_count + 1u;                              // :115
    const bool truncated = step_count >= R.n_free;                               // :116, max_steps = total_free (:459)

    // do_action (:134-166).  Whether the move bumps was worked out by the previous step's rays (record byte `pad0`), so
    // the new position — and with it every address this step touches — is known as soon as the record arrives.
    int x = st.x, y = st.y, z = st.z, facing = st.facing;
    uint32_t dir;                                                                // index into blocked6
    if (a < 4) {
        facing = (facing + a) & 3;                                               // table :135-140 == rotate by a; :148-151
        dir = (0x1302u >> (facing * 4)) & 3u;                                    // N -> +y (2), E -> +x (0), S -> -y (3), W -> -x (1)
    } else dir = (uint32_t)a;                                                    // 4 -> +z, 5 -> -z
    const bool moved = !((st.pad0 >> dir) & 1u);                                 // _mark_visited :328-332
    if (moved) {
        x += (dir == 0) - (dir == 1);
        y += (dir == 2) - (dir == 3);
        z += (dir == 4) - (dir == 5);
    }
    const bool bumped = !moved;

    // Issue every load of the step now: first window batch, the counter of the final cell, the three occupancy words.
    // (A caller that wants no observation at all — the inner steps of a fused rollout — skips the window entirely.)
    const bool any_obs = io.obs != nullptr || io.terminal_obs != nullptr;
    WinBatch<G> wb;
    if (any_obs) window_load<G>(P, R, envk, lane, x, y, z, 0, wb);
    const uint32_t cidx = c_index(R, x, y, z);
    const int c_old = C[cidx];
    const Rays r = cast_rays(P, R, x, y, z);

    // entering: 0 -> 1 (+visited, explored) or v -> v+1 (:335-341); then the unconditional += 1 at the final
    // position (:165-166).  A bump only gets the latter.
    const bool explored = moved && c_old == 0;
    const int c_new = imin(255, c_old + (moved ? 2 : 1));
    if (lane == 0) C[cidx] = (uint8_t)c_new;
    const uint32_t visited = st.visited_count + (explored ? 1u : 0u);

    // termination test of compute_reward (:212-214): visited/total >= 0.84 in f64.  For total <= 65536 this is
    // exactly 25*visited >= 21*total (tests/test_host_logic.py::test_finish_threshold_integer_form).
    const bool done = 25u * visited >= 21u * R.n_free;
    const bool will_reset = P.auto_reset && (done || truncated);

    // get_obs (:122)
    float *orow = io.obs ? io.obs + row * kObsDim : nullptr;
    if (will_reset) orow = io.terminal_obs ? io.terminal_obs + row * kObsDim : nullptr;
    if (orow != nullptr) {
        ObsScalars sc;
        sc.facing = facing; sc.last_action = st.last_action; sc.was_near_wall = (flags & kWasNearWall) != 0;
        sc.last_bump = (flags & kLastBump) != 0; sc.down = r.down; sc.visited = visited; sc.total_free = R.n_free;
        window_store<G>(R, lane, x, y, z, 0, wb, r, c_new, lut, orow);
#pragma unroll
        for (int a0 = WinMap<G>::AB; a0 < WinMap<G>::NX; a0 += WinMap<G>::AB) {
            window_load<G>(P, R, envk, lane, x, y, z, a0, wb);
            window_store<G>(R, lane, x, y, z, a0, wb, r, c_new, lut, orow);
        }
        write_scalars<G>(P, lane, sc, lut, orow);
    }
    // The cells a ray pass marks depend only on the position (L and the room are fixed), and marks are never erased
    // within an episode: every earlier stay at this cell (c_old >= 1, which includes every bump) already marked
    // them.  Only a FIRST visit has anything to write, so revisits skip the ray marking and its scattered traffic.
    if (explored && !will_reset) mark_seen<G>(R, envk, lane, x, y, z, r, (int)dir, P.L);
    if (r.near_wall) flags |= kNearWall;

    if (REG_STATE || lane == 0) {
        const bool writer = lane == 0;
        // compute_reward (:169-224), same operations in the same order, f64
        double rew = -0.05;
        const double pen = (double)c_new * 0.02;
        rew -= (pen < 0.5) ? pen : 0.5;
        int cents = -5 - imin(2 * c_new, 50);
        uint32_t bump_count = st.bump_count;
        if (bumped) {
            flags |= kLastBump; bump_count++;
            rew += P.crash_penalty;
        } else {
            flags &= ~kLastBump;
            if (flags & kWasNearWall) { flags &= ~kWasNearWall; rew += 0.15; cents += 15; }
            if (st.last_action != 2 && a == st.last_action && st.last_action < 4) { rew += 0.05; cents += 5; }
            if (st.last_action == 2 && a == 2) { rew -= 0.5; cents -= 50; }
        }
        if (explored) { rew += 1.0; cents += 100; }
        if (done) { flags |= kDone; rew += 100.0; cents += 10000; }
        if (truncated) { rew += -5.0; cents -= 500; }
        const int ret_centi = st.ret_centi + cents;

        if (writer) {
            store_stream(io.reward + row, (float)rew);
            if (io.reward64) io.reward64[row] = rew;
            if (!REG_STATE || io.terminated) io.terminated[row] = done ? 1 : 0;
            if (!REG_STATE || io.truncated) io.truncated[row] = truncated ? 1 : 0;
        }
        if (REG_STATE && done_bits) *done_bits = (done ? 1u : 0u) | (truncated ? 2u : 0u);
        if (writer && (done || truncated) && io.episodes) {
            EpisodeRec ep;
            ep.episode_return = (float)((double)ret_centi / 100.0 + (double)bump_count * P.crash_penalty);
            ep.length = (int32_t)step_count; ep.bumps = (int32_t)bump_count; ep.visited = (int32_t)visited;
            ep.total_free = (int32_t)R.n_free; ep.room = st.room; ep.terminated = done; ep.truncated = truncated;
            reinterpret_cast<EpisodeRec *>(io.episodes)[row] = ep;
        }
        if (!will_reset) {
            EnvState ns;
            ns.x = (uint8_t)x; ns.y = (uint8_t)y; ns.z = (uint8_t)z; ns.facing = (uint8_t)facing;
            ns.last_action = (uint8_t)a; ns.flags = (uint8_t)flags; ns.down = (uint8_t)r.down; ns.pad0 = (uint8_t)r.blocked6;
            ns.step_count = step_count; ns.visited_count = visited; ns.bump_count = bump_count;
            ns.ret_centi = ret_centi; ns.episode = st.episode; ns.room = st.room; ns.pad1 = 0;
            if (REG_STATE) *rs = ns;
            else P.states[env] = ns;
        }
    }
    if (INLINE_RESET) {
        if (will_reset) {
            // every lane's reads of the old knowledge are done before any lane clears it
            group_sync<G>(lane_in_warp);
            reset_env_philox<G>(P, env, lane, lane_in_warp, st.episode, lut, io.obs ? io.obs + row * kObsDim : nullptr);
        }
        return false;
    }
    return will_reset;
}

// ===============================================================================================================
// simpleEnv (envs/simpleEnv.py): ternary knowledge grid, ray-cell observations, goal reward
//   step :109-150, do_action :152-186, compute_reward :189-217, get_obs :219-265, _mark_visited :273-298,
//   _sense_direction :301-337, reset :79-107, load_room's start/goal picks :404-426.
// Knowledge: 2 bits per cell — 00 unknown (-1), 01 seen free (0), 10 visited (1), 11 marked blocked (2) — stored per (x,y)
// column as one u32 (low half = bit-plane 0, high half = bit-plane 1, bit z), in 4x4-column tiles of 64 B.
// EnvState reuse: down = goal x, pad0 = goal y, pad1 = goal z.
// ===============================================================================================================
NAV3D_HD uint32_t k2_bytes(const RoomDev &R) { return (uint32_t)R.ntx * R.nty * 64u; }
NAV3D_HD int k2_code(uint32_t w, int z) { return (int)(((w >> z) & 1u) | (((w >> (16 + z)) & 1u) << 1)); }
NAV3D_HD float k2_value(int code) { return code == 0 ? -1.0f : (float)(code - 1); }

// get_obs (simpleEnv.py:219-265) at (x,y,z): marks the knowledge and writes the 6L+7 floats.  `centre` is the centre
// column's word as every lane holds it (after the move's visit update); the updated word is stored by lane 0.
// The six rays are evaluated once per absolute axis direction a (0 +x, 1 -x, 2 +y, 3 -y, 4 +z, 5 -z) and kept packed in
// registers — free count in byte a of `nf6`, "a wall stopped it" / "a 2 is appended" in bit a of `wall6` / `blk6` — so that
// the per-cell loop below needs no indexed local array.
template <int G>
NAV3D_HD void simple_observe(const EngineParams &P, const RoomDev &R, uint32_t *K, int lane, int lane_in_warp, int x, int y,
                             int z, int facing, uint32_t centre_mem, uint32_t centre, int last_action, float *obs_row) {
    const int L = P.L, W = R.W, D = R.D, H = R.H;
    const unsigned long long wx = ldg(P.occ64 + R.occx_off + (uint32_t)(y * H + z));
    const unsigned long long wy = ldg(P.occ64 + R.occy_off + (uint32_t)(x * H + z));
    const unsigned long long wz = ldg(P.occz + R.occz_off + (uint32_t)(x * D + y));
    unsigned long long nf6 = 0;
    uint32_t wall6 = 0, blk6 = 0;
    {
        int ext, nfree, near, room_left;
#define NAV3D_SIMPLE_RAY(a, CALL, LEFT)                                                                     \
        room_left = (LEFT); CALL;                                                                           \
        nf6 |= (unsigned long long)nfree << (8 * (a));                                                      \
        wall6 |= (uint32_t)(ext > nfree) << (a);                     /* a wall stopped the ray (:321-324) */ \
        blk6 |= (uint32_t)((ext > nfree) || (nfree == room_left && nfree < L)) << (a);   /* ... or the room ended (:311-319) */
        NAV3D_SIMPLE_RAY(0, ray_up<false>(wx, x, imin(L, room_left), ext, nfree, near), W - 1 - x)
        NAV3D_SIMPLE_RAY(1, ray_down<false>(wx, x, imin(L, room_left), ext, nfree, near), x)
        NAV3D_SIMPLE_RAY(2, ray_up<false>(wy, y, imin(L, room_left), ext, nfree, near), D - 1 - y)
        NAV3D_SIMPLE_RAY(3, ray_down<false>(wy, y, imin(L, room_left), ext, nfree, near), y)
        NAV3D_SIMPLE_RAY(4, ray_up<false>(wz, z, imin(L, room_left), ext, nfree, near), H - 1 - z)
        NAV3D_SIMPLE_RAY(5, ray_down<false>(wz, z, imin(L, room_left), ext, nfree, near), z)
#undef NAV3D_SIMPLE_RAY
        (void)near;
    }
    // vertical rays: all in the centre column; fold their marks into one word
    const int nup = (int)((nf6 >> 32) & 0xff), ndn = (int)((nf6 >> 40) & 0xff);
    uint32_t free_z = (((1u << nup) - 1u) << (z + 1)) | (((1u << ndn) - 1u) << (z - ndn));
    uint32_t block_z = 0;
    if (wall6 & 16u) block_z |= 1u << (z + nup + 1);
    else if ((blk6 & 16u) && nup >= 1) block_z |= 1u << (z + nup);                 // last in-bounds cell becomes 2 (:314-317)
    if (wall6 & 32u) block_z |= 1u << (z - ndn - 1);
    else if ((blk6 & 32u) && ndn >= 1) block_z |= 1u << (z - ndn);
    const uint32_t lo0 = centre & 0xffffu, hi0 = centre >> 16;
    const uint32_t lo1 = lo0 | (free_z & ~hi0 & ~lo0);                           // unknown -> seen (:327-328)
    const uint32_t centre_seen = lo1 | (hi0 << 16);                              // what the ray cells report
    const uint32_t centre_new = (lo1 | block_z) | ((hi0 | block_z) << 16);
    // ray order :243: forward, left, right, backward, up, down.  Headings N=+y, E=+x, S=-y, W=-x; heading -> axis
    // direction index is the same nibble table the move uses (N -> 2, E -> 0, S -> 3, W -> 1).
#pragma unroll
    for (int d = 0; d < 6; d++) {
        int a, dx = 0, dy = 0, dz = 0;
        if (d < 4) {
            const int h = (facing + ((0x2130 >> (4 * d)) & 3)) & 3;              // fwd +0, left +3, right +1, back +2
            a = (0x1302 >> (4 * h)) & 3;
            dx = (h == 1) - (h == 3); dy = (h == 0) - (h == 2);
        } else { a = d; dz = (d == 4) ? 1 : -1; }
        const int nfree = (int)((nf6 >> (8 * a)) & 0xff);
        const bool wall = (wall6 >> a) & 1u, blocked = (blk6 >> a) & 1u;
        for (int s = 1 + lane; s <= L; s += G) {
            float v = -1.0f;                                                  // padding (:334-335)
            if (s <= nfree) {
                const int cx = x + dx * s, cy = y + dy * s, cz = z + dz * s;
                if (d >= 4) v = k2_value(k2_code(centre_seen, cz));
                else {
                    uint32_t *p = K + s_index(R, cx, cy);
                    uint32_t w = *p, n = w;
                    int code = k2_code(w, cz);
                    if (code == 0) { code = 1; n |= 1u << cz; }               // -1 -> 0
                    v = k2_value(code);
                    if (s == nfree && blocked && !wall) n |= (1u << cz) | (1u << (16 + cz));   // then becomes 2
                    if (n != w) *p = n;
                }
            } else if (s == nfree + 1 && blocked) {
                v = 2.0f;
                if (wall && d < 4) {
                    const int cx = x + dx * s, cy = y + dy * s;
                    uint32_t *p = K + s_index(R, cx, cy);
                    uint32_t w = *p, n = w | (1u << z) | (1u << (16 + z));
                    if (n != w) *p = n;
                }
            }
            if (obs_row) obs_row[d * L + s - 1] = v;
        }
    }
    if (obs_row) {
        for (int i = lane; i < 7; i += G) {
            float v = (float)last_action;
            if (i < 6) {
                int a = i;
                if (i < 4) { const int h = (facing + ((0x2130 >> (4 * i)) & 3)) & 3; a = (0x1302 >> (4 * h)) & 3; }
                v = ldg(P.dist_lut + (int)((nf6 >> (8 * a)) & 0xff));
            }
            obs_row[6 * L + i] = v;
        }
    }
    if (lane == 0 && centre_new != centre_mem) K[s_index(R, x, y)] = centre_new;
    (void)lane_in_warp;
}

template <int G>
NAV3D_HD void simple_reset_env(const EngineParams &P, int env, int lane, int lane_in_warp, uint32_t room_idx, uint32_t k,
                               uint32_t kg, uint32_t episode_after, float *obs_row) {
    const RoomDev R = P.rooms[room_idx];
    uint32_t *K = reinterpret_cast<uint32_t *>(P.know + (unsigned long long)env * P.env_stride);
    {
        uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        uint4 *k4 = reinterpret_cast<uint4 *>(K);
        const uint32_t n = k2_bytes(R) >> 4;
        for (uint32_t i = lane; i < n; i += G) k4[i] = zero;
    }
    group_sync<G>(lane_in_warp);
    const uint32_t cell = ldg(P.free_cells + R.free_off + k), goal = ldg(P.free_cells + R.free_off + kg);
    const int x = cell & 0xff, y = (cell >> 8) & 0xff, z = (cell >> 16) & 0xff;
    const uint32_t centre = 1u << (16 + z);                                       // internal_grid[start] = 1 (:85)
    simple_observe<G>(P, R, K, lane, lane_in_warp, x, y, z, 0, 0u, centre, 0, obs_row);
    if (lane == 0) {
        EnvState st;
        st.x = (uint8_t)x; st.y = (uint8_t)y; st.z = (uint8_t)z; st.facing = 0; st.last_action = 0; st.flags = 0;
        st.down = (uint8_t)(goal & 0xff); st.pad0 = (uint8_t)((goal >> 8) & 0xff); st.pad1 = (uint16_t)((goal >> 16) & 0xff);
        st.step_count = 0; st.visited_count = 1; st.bump_count = 0; st.ret_centi = 0;
        st.episode = episode_after; st.room = (uint16_t)room_idx;
        P.states[env] = st;
    }
}

template <int G>
NAV3D_HD void simple_reset_env_philox(const EngineParams &P, int env, int lane, int lane_in_warp, uint32_t episode,
                                      float *obs_row) {
    uint32_t u[4];
    philox4x32_10_4(P.env_id0 + (uint32_t)env, episode, 0u, kStreamReset, P.seed_lo, P.seed_hi, u);
    const uint32_t room = mulhi_range(u[0], (uint32_t)P.n_rooms);                 // random.choice(self.rooms) (:351)
    const uint32_t nf = ldg(&P.rooms[room].n_free);
    simple_reset_env<G>(P, env, lane, lane_in_warp, room, mulhi_range(u[1], nf), mulhi_range(u[2], nf), episode + 1u, obs_row);
}

template <int G>
NAV3D_HD void simple_step_env(const EngineParams &P, const StepIO &io, int env, int lane, int lane_in_warp, int action,
                              long long row) {
    const EnvState st = P.states[env];
    const RoomDev R = P.rooms[st.room];
    uint32_t *K = reinterpret_cast<uint32_t *>(P.know + (unsigned long long)env * P.env_stride);
    const int a = action < 0 ? 0 : (action > 5 ? 5 : action);
    const uint32_t step_count = st.step_count + 1u;
    const bool truncated = step_count >= R.n_free;                                 // :110-111, max_steps = total_free (:404)
    int x = st.x, y = st.y, z = st.z, facing = st.facing;
    int tx = x, ty = y, tz = z;
    if (a < 4) {
        facing = (facing + a) & 3;
        tx += (facing == 1) - (facing == 3);
        ty += (facing == 0) - (facing == 2);
    } else tz += (a == 4) ? 1 : -1;
    bool moved = false;
    if (tx >= 0 && tx < R.W && ty >= 0 && ty < R.D && tz >= 0 && tz < R.H) {
        const uint32_t ow = ldg(P.occz + R.occz_off + (uint32_t)(tx * R.D + ty));
        moved = !((ow >> tz) & 1u);
    }
    if (moved) { x = tx; y = ty; z = tz; }
    const uint32_t centre_mem = K[s_index(R, x, y)];
    uint32_t centre = centre_mem;
    group_sync<G>(lane_in_warp);                 // every lane holds the old word before lane 0 rewrites it
    bool explored = false;
    if (moved) {                                 // _mark_visited :286-294: 0 or -1 become 1; 1 and 2 stay
        const int code = k2_code(centre, z);
        if (code <= 1) { explored = true; centre = (centre & ~(1u << z)) | (1u << (16 + z)); }
    }
    const uint32_t visited = st.visited_count + (explored ? 1u : 0u);
    const int gx = st.down, gy = st.pad0, gz = st.pad1;
    bool done = (st.flags & kDone) != 0;
    int hits = 0;
    for (int i = 0; i < 5; i++) if (x == gx && y == gy && z - i == gz) hits++;     // :203-208
    done = done || hits > 0;
    const bool will_reset = P.auto_reset && (done || truncated);
    float *orow = io.obs + row * P.obs_dim;
    if (will_reset) orow = io.terminal_obs ? io.terminal_obs + row * P.obs_dim : nullptr;
    if (!will_reset || orow)
        simple_observe<G>(P, R, K, lane, lane_in_warp, x, y, z, facing, centre_mem, centre, a, orow);
    if (lane == 0) {
        double rew = -0.1;                                                         // compute_reward :191-215
        int cents = -10;
        uint32_t bump_count = st.bump_count;
        if (!moved) { bump_count++; rew += -10.0; cents -= 1000; }
        if (a != 2 && a < 4) { rew += 0.05; cents += 5; }                          // last_action was already set to a (:139)
        for (int i = 0; i < hits; i++) { rew += 100.0; cents += 10000; }
        if (explored) { rew += 1.0; cents += 100; }
        const int ret_centi = st.ret_centi + cents;
        store_stream(io.reward + row, (float)rew);
        if (io.reward64) io.reward64[row] = rew;
        io.terminated[row] = done ? 1 : 0;
        io.truncated[row] = truncated ? 1 : 0;
        if ((done || truncated) && io.episodes) {
            EpisodeRec ep;
            ep.episode_return = (float)((double)ret_centi / 100.0);
            ep.length = (int32_t)step_count; ep.bumps = (int32_t)bump_count; ep.visited = (int32_t)visited;
            ep.total_free = (int32_t)R.n_free; ep.room = st.room; ep.terminated = done; ep.truncated = truncated;
            reinterpret_cast<EpisodeRec *>(io.episodes)[row] = ep;
        }
        if (!will_reset) {
            EnvState ns = st;
            ns.x = (uint8_t)x; ns.y = (uint8_t)y; ns.z = (uint8_t)z; ns.facing = (uint8_t)facing; ns.last_action = (uint8_t)a;
            ns.flags = (uint8_t)(done ? kDone : 0u);
            ns.step_count = step_count; ns.visited_count = visited; ns.bump_count = bump_count; ns.ret_centi = ret_centi;
            P.states[env] = ns;
        }
    }
    if (will_reset) {
        group_sync<G>(lane_in_warp);
        simple_reset_env_philox<G>(P, env, lane, lane_in_warp, st.episode, io.obs + row * P.obs_dim);
    }
}

}  // namespace nav3d
