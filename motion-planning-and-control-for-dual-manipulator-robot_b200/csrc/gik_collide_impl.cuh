// gik_collide_impl.cuh -- collision kernels + their C ABI (included at the end of gik_kernels.cu: one translation
// unit, one library).  Mapping: ONE CONFIGURATION PER WARP.  The kinematic tree is walked level by level (one joint per lane),
// the lanes place the geometries (oMg = oMi[parent] * placement, pin.updateGeometryPlacements), then the collision
// pairs are dealt round-robin to the 32 lanes for the bounding-sphere test, the survivors are compacted into a per-warp
// list and boolean GJK runs on full warps of candidates, with one __any_sync per round for the early exit.  The scene (geometries, pair lists) lives in global memory and is read
// through the read-only path; per-configuration traffic is q in (60 B) and one flag out.
#pragma once
#include "gik_collide.cuh"

namespace gik {

constexpr int kCollideWarps = 4;
constexpr int kMaxCand = 512;       // per-configuration candidate pairs kept for the narrow phase

template <typename T>
__global__ void __launch_bounds__(kCollideWarps * 32)
gik_collision_kernel(const DevScene<T>* __restrict__ sc, const uint8_t* __restrict__ pa, const uint8_t* __restrict__ pb,
                     int n_pairs, int64_t n, const T* __restrict__ q, const T* __restrict__ cube_pose, T margin,
                     uint8_t* __restrict__ out, int invert, const int64_t* __restrict__ sel, int64_t n_sel,
                     const int64_t* __restrict__ n_sel_dev) {
  __shared__ T s_oMi[kCollideWarps][GIK_MAX_NQ][12];
  __shared__ T s_oMg[kCollideWarps][GIK_MAX_GEOMS][12];
  __shared__ uint16_t s_cand[kCollideWarps][kMaxCand];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nq = sc->tree.nq, ng = sc->n_geoms;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // `sel` (optional): test only the columns sel[0 .. n_sel) of the [.][n] arrays; results go to out[sel[.]]
  const int64_t count = sel ? (n_sel_dev ? *n_sel_dev : n_sel) : n;
  for (int64_t w_i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w_i < count; w_i += warps) {
    const int64_t i = sel ? __ldg(sel + w_i) : w_i;
    if (q) {
      // tree FK, level by level: lane j first builds its joint's local transform jointPlacement * Rot(axis, q_j), then
      // the joints of depth d (all independent) are composed with their parents' world placements of depth d - 1
      T L[12];
      int par = -1, dep = -1;
      if (lane < nq) {
        joint_placement(sc->tree, lane, (const T*)nullptr, __ldg(q + (int64_t)lane * n + i), L);
        par = sc->tree.parent[lane]; dep = sc->tree.depth[lane];
      }
      for (int d = 0; d <= sc->tree.max_depth; ++d) {
        if (dep == d) {
          if (par < 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) s_oMi[w][lane][k] = L[k];
          } else {
            se3_mul12(&s_oMi[w][par][0], L, &s_oMi[w][lane][0]);
          }
        }
        __syncwarp();
      }
    }
    for (int g = lane; g < ng; g += 32) {
      const DevGeom<T>& G = sc->g[g];
      T P[12];
      if (g == sc->cube_geom && cube_pose) {
#pragma unroll
        for (int k = 0; k < 12; ++k) P[k] = __ldg(cube_pose + (int64_t)k * n + i);
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) P[k] = G.R[k];
        P[9] = G.p[0]; P[10] = G.p[1]; P[11] = G.p[2];
      }
      if (G.joint >= 0 && q) se3_mul12(&s_oMi[w][G.joint][0], P, &s_oMg[w][g][0]);
      else {
#pragma unroll
        for (int k = 0; k < 12; ++k) s_oMg[w][g][k] = P[k];
      }
    }
    __syncwarp();
    // broad phase over all pairs first (bounding spheres), compacting the survivors into a per-warp list, so that the
    // narrow phase (GJK, a few hundred instructions, data dependent) runs on FULL warps of candidates instead of on
    // the one or two lanes of each 32-pair round that happen to survive
    bool hit = false;
    int n_cand = 0;
    for (int k0 = 0; k0 < n_pairs && !hit; k0 += 32) {
      const int k = k0 + lane;
      bool pass = false;
      if (k < n_pairs) {
        const int a = pa[k], b = pb[k];
        const T* Ma = &s_oMg[w][a][0];
        const T* Mb = &s_oMg[w][b][0];
        const T dx = Ma[9] - Mb[9], dy = Ma[10] - Mb[10], dz = Ma[11] - Mb[11];
        const T reach = sc->g[a].bound + sc->g[b].bound + margin;
        pass = dx * dx + dy * dy + dz * dz <= reach * reach;
      }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (n_cand + __popc(m) > kMaxCand) {
        // list full (never with the reference scene): test this round's survivors in place
        if (pass) hit = gjk_intersect(make_shape(sc->g[pa[k]], &s_oMg[w][pa[k]][0]), make_shape(sc->g[pb[k]], &s_oMg[w][pb[k]][0]), margin);
        hit = __any_sync(0xffffffffu, hit);
      } else {
        if (pass) s_cand[w][n_cand + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
        n_cand += __popc(m);
      }
    }
    __syncwarp();
    for (int c0 = 0; c0 < n_cand && !hit; c0 += 32) {
      if (c0 + lane < n_cand) {
        const int k = s_cand[w][c0 + lane];
        const int a = pa[k], b = pb[k];
        hit = gjk_intersect(make_shape(sc->g[a], &s_oMg[w][a][0]), make_shape(sc->g[b], &s_oMg[w][b][0]), margin);
      }
      hit = __any_sync(0xffffffffu, hit);
    }
    if (lane == 0) out[i] = (uint8_t)((hit ? 1 : 0) ^ (invert ? 1 : 0));
    __syncwarp();
  }
}

// converged problems -> index list (order irrelevant) + count, everything else gets success = 0; one atomicAdd per warp
__global__ void __launch_bounds__(256) gik_compact_converged_kernel(int64_t n, const uint8_t* __restrict__ conv,
                                                                    uint8_t* __restrict__ success, int64_t* __restrict__ count,
                                                                    int64_t* __restrict__ sel) {
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + lane;
    const bool c = i < n && conv[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, c);
    long long base = 0;
    if (lane == 0 && m) base = (long long)atomicAdd((unsigned long long*)count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (c) sel[base + __popc(m & ((1u << lane) - 1u))] = i;
    else if (i < n) success[i] = 0;
  }
}

}  // namespace gik

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct gik_scene_dev {
  gik::DevScene<float>* sc32 = nullptr;
  gik::DevScene<double>* sc64 = nullptr;
  uint8_t* pairs = nullptr;   // [3 lists][2][GIK_MAX_PAIRS]: all pairs | table/obstacle pairs | cube vs table, obstacle
  int n_list[3] = {0, 0, 0};
};

namespace {

void free_scene(gik_scene_dev* sd) {
  if (!sd) return;
  cudaFree(sd->sc32); cudaFree(sd->sc64); cudaFree(sd->pairs);
  delete sd;
}

template <typename T> gik::DevScene<T>* scene_of(gik_scene_dev* sd);
template <> gik::DevScene<float>* scene_of<float>(gik_scene_dev* sd) { return sd->sc32; }
template <> gik::DevScene<double>* scene_of<double>(gik_scene_dev* sd) { return sd->sc64; }

// list: 0 all pairs, 1 table/obstacle pairs, 2 cube's own pairs
template <typename T>
int collide_api(gik_handle_t h, int64_t n, const T* q, const T* cube_pose, int list, double margin, int invert,
                uint8_t* out, void* stream, const int64_t* sel = nullptr, int64_t n_sel = 0,
                const int64_t* n_sel_dev = nullptr) {
  if (bad_handle(h)) return GIK_E_HANDLE;
  if (n < 0) return GIK_E_SIZE;
  if (!(margin >= 0.0)) return GIK_E_PARAM;
  if (!h->scene) return GIK_E_NOSCENE;
  if (n == 0 || (sel && !n_sel_dev && n_sel == 0)) return GIK_OK;
  if (n_sel < 0) return GIK_E_SIZE;
  if (!out || (list != 2 && !q) || (list == 2 && !cube_pose)) return GIK_E_NULL;
  DeviceGuard g(h->device);
  if (g.err != cudaSuccess) return (int)g.err;
  gik_scene_dev* sd = h->scene;
  const int64_t work = (sel && !n_sel_dev) ? n_sel : n;     // device-side count: size the grid for the upper bound
  int64_t blocks = (work + gik::kCollideWarps - 1) / gik::kCollideWarps;
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  const uint8_t* pa = sd->pairs + (size_t)list * 2 * GIK_MAX_PAIRS;
  gik::gik_collision_kernel<T><<<(int)blocks, gik::kCollideWarps * 32, 0, (cudaStream_t)stream>>>(
      scene_of<T>(sd), pa, pa + GIK_MAX_PAIRS, sd->n_list[list], n, q, cube_pose, (T)margin, out, invert, sel, n_sel, n_sel_dev);
  return (int)cudaGetLastError();
}

}  // namespace

// K3 + collision in one stream-ordered call, no host synchronisation: solve, compact the converged problems on the
// device, test exactly those (the predicate's short-circuit) and write success = converged && !colliding.
template <typename T>
int solve_success_api(gik_handle_t h, int64_t n, const T* q_init, const T* pose, const gik_params_t* prm, T* q_out,
                      uint8_t* success, uint8_t* conv, int32_t* iters, T* resid, int64_t* scratch, void* stream) {
  if (bad_handle(h)) return GIK_E_HANDLE;
  if (!h->scene) return GIK_E_NOSCENE;
  if (n > 0 && (!success || !conv || !scratch)) return GIK_E_NULL;
  int rc = solve_api<T>(h, n, q_init, pose, prm, q_out, conv, iters, resid, stream);
  if (rc || n == 0) return rc;
  DeviceGuard g(h->device);
  if (g.err != cudaSuccess) return (int)g.err;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int64_t), st);
  if (e != cudaSuccess) return (int)e;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
  gik::gik_compact_converged_kernel<<<(int)blocks, 256, 0, st>>>(n, conv, success, scratch, scratch + 1);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  return collide_api<T>(h, n, q_out, pose, 0, 0.0, 1, success, stream, scratch + 1, 0, scratch);
}

extern "C" {

int gik_scene_attach(gik_handle_t h, const gik_scene_t* scene) {
  if (bad_handle(h)) return GIK_E_HANDLE;
  if (!scene) return GIK_E_NULL;
  int rc = gik::validate_scene(h->host, *scene);
  if (rc) return rc;
  DeviceGuard g(h->device);
  if (g.err != cudaSuccess) return (int)g.err;
  gik_scene_dev* sd = new (std::nothrow) gik_scene_dev();
  if (!sd) return (int)cudaErrorMemoryAllocation;
  gik::DevScene<float>* h32 = new (std::nothrow) gik::DevScene<float>();
  gik::DevScene<double>* h64 = new (std::nothrow) gik::DevScene<double>();
  uint8_t* hp = new (std::nothrow) uint8_t[3 * 2 * GIK_MAX_PAIRS]();
  cudaError_t e = (h32 && h64 && hp) ? cudaSuccess : cudaErrorMemoryAllocation;
  if (e == cudaSuccess) {
    gik::fill_dev_scene(h->host, *scene, *h32);
    gik::fill_dev_scene(h->host, *scene, *h64);
    for (int k = 0; k < scene->n_pairs; ++k) {
      const uint8_t a = scene->pair_a[k], b = scene->pair_b[k];
      hp[sd->n_list[0]] = a; hp[GIK_MAX_PAIRS + sd->n_list[0]++] = b;
      if ((int)b == scene->table_geom || (int)b == scene->obstacle_geom) {          // tools.py:42
        hp[2 * GIK_MAX_PAIRS + sd->n_list[1]] = a; hp[3 * GIK_MAX_PAIRS + sd->n_list[1]++] = b;
      }
    }
    const int env[2] = {scene->table_geom, scene->obstacle_geom};                    // setup_pinocchio.py:66-70
    for (int k = 0; k < 2; ++k)
      if (env[k] >= 0 && scene->cube_geom >= 0) {
        hp[4 * GIK_MAX_PAIRS + sd->n_list[2]] = (uint8_t)env[k]; hp[5 * GIK_MAX_PAIRS + sd->n_list[2]++] = (uint8_t)scene->cube_geom;
      }
    if ((e = cudaMalloc(&sd->sc32, sizeof(*h32))) == cudaSuccess &&
        (e = cudaMalloc(&sd->sc64, sizeof(*h64))) == cudaSuccess &&
        (e = cudaMalloc(&sd->pairs, 3 * 2 * GIK_MAX_PAIRS)) == cudaSuccess &&
        (e = cudaMemcpy(sd->sc32, h32, sizeof(*h32), cudaMemcpyHostToDevice)) == cudaSuccess &&
        (e = cudaMemcpy(sd->sc64, h64, sizeof(*h64), cudaMemcpyHostToDevice)) == cudaSuccess)
      e = cudaMemcpy(sd->pairs, hp, 3 * 2 * GIK_MAX_PAIRS, cudaMemcpyHostToDevice);
  }
  delete h32; delete h64; delete[] hp;
  if (e != cudaSuccess) { free_scene(sd); return (int)e; }
  cudaDeviceSynchronize();          // a previous scene may still be in use by enqueued kernels
  free_scene(h->scene);
  h->scene = sd;
  return GIK_OK;
}

int gik_collision_f32(gik_handle_t h, int64_t n, const float* q, const float* cube, uint8_t* out, void* s) { return collide_api<float>(h, n, q, cube, 0, 0.0, 0, out, s); }
int gik_collision_f64(gik_handle_t h, int64_t n, const double* q, const double* cube, uint8_t* out, void* s) { return collide_api<double>(h, n, q, cube, 0, 0.0, 0, out, s); }
int gik_collision_sel_f32(gik_handle_t h, int64_t n, int64_t n_sel, const int64_t* sel, const float* q, const float* cube, uint8_t* out, void* s) {
  if (!sel) return GIK_E_NULL;
  return collide_api<float>(h, n, q, cube, 0, 0.0, 0, out, s, sel, n_sel);
}
int gik_collision_sel_f64(gik_handle_t h, int64_t n, int64_t n_sel, const int64_t* sel, const double* q, const double* cube, uint8_t* out, void* s) {
  if (!sel) return GIK_E_NULL;
  return collide_api<double>(h, n, q, cube, 0, 0.0, 0, out, s, sel, n_sel);
}
int gik_solve_success_f32(gik_handle_t h, int64_t n, const float* q_init, const float* pose, const gik_params_t* p, float* q_out,
                          uint8_t* success, uint8_t* conv, int32_t* iters, float* resid, int64_t* scratch, void* s) {
  return solve_success_api<float>(h, n, q_init, pose, p, q_out, success, conv, iters, resid, scratch, s);
}
int gik_solve_success_f64(gik_handle_t h, int64_t n, const double* q_init, const double* pose, const gik_params_t* p, double* q_out,
                          uint8_t* success, uint8_t* conv, int32_t* iters, double* resid, int64_t* scratch, void* s) {
  return solve_success_api<double>(h, n, q_init, pose, p, q_out, success, conv, iters, resid, scratch, s);
}
int gik_clearance_f32(gik_handle_t h, int64_t n, const float* q, const float* cube, double thr, uint8_t* out, void* s) { return collide_api<float>(h, n, q, cube, 1, thr, 1, out, s); }
int gik_clearance_f64(gik_handle_t h, int64_t n, const double* q, const double* cube, double thr, uint8_t* out, void* s) { return collide_api<double>(h, n, q, cube, 1, thr, 1, out, s); }
int gik_cube_collision_f32(gik_handle_t h, int64_t n, const float* cube, uint8_t* out, void* s) { return collide_api<float>(h, n, (const float*)nullptr, cube, 2, 0.0, 0, out, s); }
int gik_cube_collision_f64(gik_handle_t h, int64_t n, const double* cube, uint8_t* out, void* s) { return collide_api<double>(h, n, (const double*)nullptr, cube, 2, 0.0, 0, out, s); }

}  // extern "C"
