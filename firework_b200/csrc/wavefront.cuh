// Wavefront path tracer kernels: raygen -> [extend -> miss / shade_<material>]* -> accumulate -> resolve.
// Iterative form of render.rs:12-33 `color()` (depth cap 10) inside render.rs:163-196 `render_pixel`.
//
// Path p of a batch covers pixel  pix0 + p % npix  and sample  s0 + p / npix  (implicit mapping: no id arrays).
// The queues carry the path's working state with them (ray, winning hit) as SoA float4 streams in queue order,
// compacted with warp ballot / popc, so every kernel reads its input fully coalesced and without a dependent
// "queue entry -> path -> record" hop; only the attenuation chain and the final radiance are indexed by path.
// extend publishes just the winning (t, object, primitive, barycentrics); the shade kernel that consumes a path
// rebuilds point / normal / uv (finalize_hit) in registers.
#pragma once
#include "intersect.cuh"
#include "shade.cuh"

namespace fw {

#ifndef FW_SHADE_MIN_BLOCKS
#define FW_SHADE_MIN_BLOCKS 8    // shade kernels wait on dependent loads: cap registers at 64 (10 blocks = 51 regs spills: measured slower)
#endif
#ifndef FW_EXTEND_MIN_BLOCKS
#define FW_EXTEND_MIN_BLOCKS 8   // __launch_bounds__ min blocks / SM of the BVH extend kernels (register cap knob)
#endif
constexpr int FW_MAX_DEPTH = 10;          // render.rs:21  `depth < 10`
constexpr int FW_NUM_QUEUES = 8;          // [0..5] material queues (MatKind), [6] next extend queue, [7] mesh queue (two-pass extend)
constexpr int FW_Q_EXTEND = 6, FW_Q_MESH = 7;
constexpr int FW_TILE = 128;              // paths per tile: bounce 0 deals tiles round-robin to the segments
#ifndef FW_BLOCK_THREADS
#define FW_BLOCK_THREADS 128
#endif
constexpr int FW_BLOCK = FW_BLOCK_THREADS;             // threads per block of every queue-driven kernel

// Queues are SEGMENTED: each of the `nseg` segments owns a fixed region of `seg_cap` slots in every queue and is
// processed by exactly one thread block per kernel, which is also the only writer of that segment's regions in
// the kernel's output queues.  A segment's paths therefore stay in the segment for the whole batch (its
// population only shrinks, so `seg_cap` = its bounce-0 share always suffices), fill counts are plain
// shared-memory counters written back once per block, and no global atomic is issued anywhere.  (One global
// counter per queue was the top stall of both extend and shade: ~2.5 same-sector atomics per warp iteration.)
// Path-state streams are touched once per kernel: load / store them with the streaming (evict-first) policy so
// that they do not push the scene tables (nodes, objects, shapes, materials) out of L1.
#ifndef FW_STREAM_HINTS
#define FW_STREAM_HINTS 1
#endif
FW_DEV float4 ld_stream(const float4* p) { return FW_STREAM_HINTS ? __ldcs(p) : *p; }
#ifndef FW_PREFETCH
#define FW_PREFETCH 1
#endif
// The record a shade thread will need in its NEXT iteration is FW_BLOCK slots ahead: ask L2 for it now (no register
// cost).  Measured: +1-2 % on the latency-bound shade kernels, nothing on the issue-bound extend kernels (not used there).
FW_DEV void prefetch_l2(const void* p) {
    if (FW_PREFETCH) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
FW_DEV void st_stream(float4* p, float4 v) { if (FW_STREAM_HINTS) __stcs(p, v); else *p = v; }

struct HitQueue {      // one shade queue: records of the paths whose ray hit a surface of that material
    float4* o;         // [nseg][seg_cap] ray origin.xyz, asfloat(path)
    float4* d;         // ray direction.xyz (never normalised: ray.rs), winning t        (miss queue: d.xyz, asfloat(path))
    float4* w;         // asfloat(object), asfloat(primitive), asfloat(material), asfloat(rank)   (rank: two-pass extend only)
    float4* b;         // triangle barycentrics b0, b1, b2 — allocated only for scenes with TriangleMesh objects
};
struct PathState {
    float4* xo[2];     // ping-pong extend queues: ray origin.xyz, asfloat(path)
    float4* xd[2];     //                          ray direction.xyz
    HitQueue hq[8];    // [0..5] per-material shade queues (MatKind; MAT_MISS uses .d only), [7] mesh queue (two-pass extend)
    float4* atten;     // [FW_MAX_DEPTH][cap] attenuation chain, by path (see fold_radiance)
    float4* radiance;  // [cap] finished path radiance, by path
    uint32_t* counters;                // [FW_MAX_DEPTH + 2][FW_NUM_QUEUES][nseg] fill counts
    uint32_t cap;
    uint32_t nseg, seg_cap;
};
// Counter rows: material / mesh queues of bounce b live in row b; the extend queue CONSUMED at bounce b lives in
// row b (raygen fills row 0, the shade kernels of bounce b fill row b + 1).
FW_DEV uint32_t* counter_row(const PathState& ps, uint32_t row, int queue) {
    return ps.counters + ((size_t)row * FW_NUM_QUEUES + queue) * ps.nseg;
}

struct Batch {
    uint32_t pix0, npix, s0, ns;
    uint32_t width, height;
    uint32_t npix_magic;   // ceil(2^32 / npix) (0xffffffff for npix == 1): p / npix without the emulated 32-bit division
};

// pixel = pix0 + p % npix, sample = s0 + p / npix.  q = umulhi(p, ceil(2^32 / npix)) is floor(p / npix) or one more
// (one less for npix == 1); one correction step makes it exact for every 32-bit p.
FW_DEV void batch_path(const Batch& b, uint32_t p, uint32_t& pixel, uint32_t& sample) {
    uint32_t q = __umulhi(p, b.npix_magic);
    int r = (int)(p - q * b.npix);
    if (r < 0) { q -= 1u; r += (int)b.npix; }
    else if (r >= (int)b.npix) { q += 1u; r -= (int)b.npix; }
    pixel = b.pix0 + (uint32_t)r;
    sample = b.s0 + q;
}

// render.rs:173-180 + camera.rs:109-116 + util.rs:31-33
FW_DEV void primary_ray(const CameraRec& cam, uint32_t width, uint32_t height, uint32_t pixel, uint32_t sample,
                        uint2 seed, float3& o, float3& d) {
    uint32_t px = pixel % width;
    uint32_t py = height - pixel / width;  // Coord::from_index: y counts down from `height`
    RngKey key{seed, pixel, sample, 0u};
    PhiloxStream rng(key, STREAM_CAMERA);
    float u = ((float)px + rng.next()) / (float)width;
    float v = ((float)py + rng.next()) / (float)height;
    float3 rd = cam.lens_radius * random_in_unit_disk(rng);
    float3 cu = f3(cam.u[0], cam.u[1], cam.u[2]), cv = f3(cam.v[0], cam.v[1], cam.v[2]);
    float3 pos = f3(cam.position[0], cam.position[1], cam.position[2]);
    float3 ll = f3(cam.lower_left[0], cam.lower_left[1], cam.lower_left[2]);
    float3 hor = f3(cam.horizontal[0], cam.horizontal[1], cam.horizontal[2]);
    float3 ver = f3(cam.vertical[0], cam.vertical[1], cam.vertical[2]);
    float3 offset = cu * rd.x + cv * rd.y;
    o = pos + offset;
    d = ll + u * hor + v * ver - pos - offset;
}

// One block per segment.  Tile t (FW_TILE consecutive paths) belongs to segment t % nseg, so every segment gets an
// even mix of the image; entry e of segment `seg` is path ((e / TILE) * nseg + seg) * TILE + e % TILE.  The rays
// go straight into the segment's bounce-0 extend queue.
__global__ void __launch_bounds__(FW_BLOCK) raygen_kernel(CameraRec cam, Batch b, uint2 seed, PathState ps) {
    const uint32_t total = b.npix * b.ns;
    const uint32_t seg = blockIdx.x;
    const size_t base = (size_t)seg * ps.seg_cap;
    uint32_t count = 0;
    for (uint32_t e = threadIdx.x; e < ps.seg_cap; e += FW_BLOCK) {
        uint32_t p = ((e / FW_TILE) * ps.nseg + seg) * FW_TILE + (e % FW_TILE);
        if (p >= total) break;   // p grows with e: the valid entries are a prefix
        uint32_t pixel, sample;
        batch_path(b, p, pixel, sample);
        float3 o, d;
        primary_ray(cam, b.width, b.height, pixel, sample, seed, o, d);
        st_stream(&ps.xo[0][base + e], make_float4(o.x, o.y, o.z, __uint_as_float(p)));
        st_stream(&ps.xd[0][base + e], make_float4(d.x, d.y, d.z, 0.0f));
        st_stream(&ps.radiance[p], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        count = e + 1;
    }
    // the segment's entry count = 1 + the largest valid e over the block
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (count) atomicMax(&s_count, count);
    __syncthreads();
    if (threadIdx.x == 0) counter_row(ps, 0, FW_Q_EXTEND)[seg] = s_count;
}

// ---- segment plumbing ------------------------------------------------------------------------------------
// Every queue-driven kernel runs one block per segment (gridDim.x == nseg).  Fill counters of the block's segment
// live in shared memory while the block runs.  `open` continues from the counts already in global memory: earlier
// kernels of the same bounce append to the same regions (the shade kernels share the next extend queue, pass 2
// continues pass 1's material queues); rows nobody wrote yet hold the zeros of the per-batch memset.
template <int NQ>
FW_DEV void seg_open(uint32_t* s_fill, const PathState& ps, const uint32_t* counter_rows /* [NQ][nseg] */, uint32_t seg) {
    if (threadIdx.x < NQ) s_fill[threadIdx.x] = counter_rows[(size_t)threadIdx.x * ps.nseg + seg];
    __syncthreads();
}
template <int NQ>
FW_DEV void seg_close(const uint32_t* s_fill, const PathState& ps, uint32_t* counter_rows, uint32_t seg) {
    __syncthreads();
    if (threadIdx.x < NQ) counter_rows[(size_t)threadIdx.x * ps.nseg + seg] = s_fill[threadIdx.x];
}
// Reserve one slot in queue `mine` of the segment for every lane with 0 <= mine < NQ: one shared-memory atomic per
// warp per non-empty queue (warp ballot / popc compaction).  Returns the slot (global index); all 32 lanes must call.
template <int NQ>
FW_DEV uint32_t seg_reserve(uint32_t* s_fill, uint32_t seg_base, int mine) {
    unsigned lane = threadIdx.x & 31u;
    unsigned lt = (1u << lane) - 1u;
    uint32_t slot = 0;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        unsigned mask = __ballot_sync(0xffffffffu, mine == k);
        if (mask == 0u) continue;
        int leader = __ffs(mask) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(&s_fill[k], (uint32_t)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mine == k) slot = seg_base + base + __popc(mask & lt);
    }
    return slot;
}

// extend: closest hit for every queued ray; the ray moves on, together with the winning (t, object, primitive,
// barycentrics), into its material's shade queue (or the miss queue).  The full hit record is rebuilt by the
// shade kernel that consumes it (finalize_hit), so it never travels through HBM.

// The shade queue a winner belongs to, and its material (which travels with the record so that the shade kernel can
// fetch its material in parallel with the object records).
FW_DEV int classify_winner(const DeviceScene& S, const Winner& w, int& material) {
    material = -1;
    if (!w.found) return MAT_MISS;
    int4 meta = __ldg(&S.obj_meta[w.obj]);
    material = meta.y;
    if ((meta.x & OBJ_KIND_MASK) == SH_RECT3D) material = winner_material(S, w.obj, w.h.prim);  // faces may carry their own material
    return __ldg(&S.mats[material].kind);
}
// Writes the hit record of one ray into slot `slot` of queue K (K is a compile-time constant at every call site, so
// the queue's pointers are plain kernel parameters).
template <int K>
FW_DEV void put_hit(const PathState& ps, uint32_t slot, float3 o, float3 d, uint32_t path, const Winner& w, int material) {
    if (K == MAT_MISS) {
        st_stream(&ps.hq[K].d[slot], make_float4(d.x, d.y, d.z, __uint_as_float(path)));
        return;
    }
    st_stream(&ps.hq[K].o[slot], make_float4(o.x, o.y, o.z, __uint_as_float(path)));
    st_stream(&ps.hq[K].d[slot], make_float4(d.x, d.y, d.z, w.t));
    st_stream(&ps.hq[K].w[slot], make_float4(__int_as_float(w.obj), __int_as_float(w.h.prim), __int_as_float(material), __int_as_float(w.rank)));
    if (ps.hq[K].b) st_stream(&ps.hq[K].b[slot], make_float4(w.h.b0, w.h.b1, w.h.b2, 0.0f));
}
template <int NQ>
FW_DEV void enqueue_hit(const PathState& ps, uint32_t* s_fill, uint32_t seg_base, int mine, float3 o, float3 d, uint32_t path,
                        const Winner& w, int material) {
    uint32_t slot = seg_reserve<NQ>(s_fill, seg_base, mine);
    switch (mine) {   // one arm per queue keeps the queue index a compile-time constant
        case 0: put_hit<0>(ps, slot, o, d, path, w, material); break;
        case 1: put_hit<1>(ps, slot, o, d, path, w, material); break;
        case 2: put_hit<2>(ps, slot, o, d, path, w, material); break;
        case 3: put_hit<3>(ps, slot, o, d, path, w, material); break;
        case 4: put_hit<4>(ps, slot, o, d, path, w, material); break;
        case 5: put_hit<5>(ps, slot, o, d, path, w, material); break;
        case FW_Q_MESH: if (NQ > FW_Q_MESH) put_hit<FW_Q_MESH>(ps, slot, o, d, path, w, material); break;
        default: break;
    }
}
// A hit record read back by the kernel that consumes queue K.
struct HitIn {
    float3 o, d;
    uint32_t path;
    int material;
    Winner w;
};
template <int K>
FW_DEV HitIn get_hit(const PathState& ps, uint32_t slot) {
    float4 A = ld_stream(&ps.hq[K].o[slot]), B = ld_stream(&ps.hq[K].d[slot]), C = ld_stream(&ps.hq[K].w[slot]);
    HitIn h;
    h.o = f3(A); h.d = f3(B);
    h.path = __float_as_uint(A.w);
    h.material = __float_as_int(C.z);
    h.w.found = __float_as_int(C.x) >= 0;
    h.w.t = B.w; h.w.obj = __float_as_int(C.x); h.w.rank = __float_as_int(C.w);
    h.w.h.t = B.w; h.w.h.prim = __float_as_int(C.y);
    h.w.h.b0 = h.w.h.b1 = h.w.h.b2 = 0.0f;
    if (ps.hq[K].b) {   // scenes with meshes: barycentrics of the winning triangle (unused for other shapes)
        float4 D = ld_stream(&ps.hq[K].b[slot]);
        h.w.h.b0 = D.x; h.w.h.b1 = D.y; h.w.h.b2 = D.z;
    }
    return h;
}

// Common shape of the extend kernels: one block per segment, FW_BLOCK rays per iteration.  The body sets `w`
// for valid lanes; the epilogue sorts the ray into its shade queue.
#define FW_EXTEND_PROLOGUE(NQ_OUT)                                                                         \
    __shared__ uint32_t s_fill[FW_NUM_QUEUES];                                                               \
    const uint32_t in_count = counter_row(ps, bounce, FW_Q_EXTEND)[blockIdx.x];                              \
    seg_open<NQ_OUT>(s_fill, ps, counter_row(ps, bounce, 0), blockIdx.x);                                    \
    for (uint32_t e0 = 0; e0 < in_count; e0 += FW_BLOCK) {                                                   \
        const bool valid = e0 + threadIdx.x < in_count;                                                      \
        uint32_t path = 0;                                                                                   \
        float3 o = f3(1e30f, 1e30f, 1e30f), d = f3(1.0f, 1.0f, 1.0f);                                        \
        if (valid) {                                                                                         \
            const size_t slot_in = (size_t)blockIdx.x * ps.seg_cap + e0 + threadIdx.x;                       \
            float4 ro = ld_stream(&ps.xo[bounce & 1][slot_in]), rd = ld_stream(&ps.xd[bounce & 1][slot_in]);     \
            o = f3(ro); d = f3(rd); path = __float_as_uint(ro.w);                                            \
        }                                                                                                    \
        Winner w;                                                                                            \
        w.found = false; w.t = 0.0f; w.obj = -1; w.rank = -1; w.h.t = 0.0f; w.h.prim = 0;                    \
        w.h.b0 = w.h.b1 = w.h.b2 = 0.0f;                                                                     \
        int mine = -1, material = -1;
#define FW_EXTEND_EPILOGUE(NQ_OUT)                                                                         \
        if (valid && mine < 0) mine = classify_winner(S, w, material);                                       \
        enqueue_hit<NQ_OUT>(ps, s_fill, blockIdx.x * ps.seg_cap, mine, o, d, path, w, material);             \
    }                                                                                                        \
    seg_close<NQ_OUT>(s_fill, ps, counter_row(ps, bounce, 0), blockIdx.x);

// Debug twin of the BVH extend: also records the number of box tests each path needed (FW_DEBUG_STEPS=1).
__global__ void __launch_bounds__(FW_BLOCK) extend_bvh_debug_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed, uint32_t bounce,
                                                                    uint32_t* steps) {
    FW_EXTEND_PROLOGUE(MAT_NUM_QUEUES)
        if (valid) {
            RngKey key{seed, 0u, 0u, bounce};
            batch_path(b, path, key.pixel, key.sample);
            Counters cnt{0, 0};
            trace_unified<true, true>(S, o, d, key, w, &cnt);
            steps[path] = (uint32_t)cnt.node_tests;
        }
    FW_EXTEND_EPILOGUE(MAT_NUM_QUEUES)
}

// Two-pass extend for BVH scenes with TriangleMesh objects (see UnifiedWalker PHASE).  Pass 1 settles every ray
// against the non-mesh objects and the mesh root boxes; rays that must enter a mesh are compacted into the mesh
// queue (with the pass-1 winner and its rank) and finished by pass 2, where every lane of a warp is doing real
// mesh traversal.
template <bool NESTED>
__global__ void __launch_bounds__(FW_BLOCK, FW_EXTEND_MIN_BLOCKS) extend_pass1_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed,
                                                                                uint32_t bounce) {
    FW_EXTEND_PROLOGUE(FW_NUM_QUEUES)
        if (valid) {
            RngKey key{seed, 0u, 0u, bounce};
            batch_path(b, path, key.pixel, key.sample);
            UnifiedWalker<false, NESTED, true, 1> wk;
            int stack_code[FW_STACK];
            float stack_te[FW_STACK];
            if (wk.init(S, o, d, stack_code, stack_te, nullptr)) {
                while (wk.step(S, key, nullptr)) {
                }
            }
            w = wk.w;
            if (wk.pending) mine = FW_Q_MESH;   // queue 6 is never selected (its counter slot belongs to the shade kernels)
        }
    FW_EXTEND_EPILOGUE(FW_NUM_QUEUES)
}
template <bool NESTED>
__global__ void __launch_bounds__(FW_BLOCK, FW_EXTEND_MIN_BLOCKS) extend_pass2_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed,
                                                                                uint32_t bounce) {
    __shared__ uint32_t s_fill[FW_NUM_QUEUES];
    const uint32_t in_count = counter_row(ps, bounce, FW_Q_MESH)[blockIdx.x];
    if (in_count == 0) return;  // block-uniform
    seg_open<MAT_NUM_QUEUES>(s_fill, ps, counter_row(ps, bounce, 0), blockIdx.x);   // continues pass 1's material queues
    for (uint32_t e0 = 0; e0 < in_count; e0 += FW_BLOCK) {
        const bool valid = e0 + threadIdx.x < in_count;
        int mine = -1, material = -1;
        HitIn h;
        h.path = 0; h.o = h.d = f3(0.0f, 0.0f, 0.0f);
        h.w.found = false; h.w.t = 0.0f; h.w.obj = -1; h.w.rank = -1; h.w.h.t = 0.0f; h.w.h.prim = 0;
        h.w.h.b0 = h.w.h.b1 = h.w.h.b2 = 0.0f;
        if (valid) {
            h = get_hit<FW_Q_MESH>(ps, blockIdx.x * ps.seg_cap + e0 + threadIdx.x);
            RngKey key{seed, 0u, 0u, bounce};
            batch_path(b, h.path, key.pixel, key.sample);
            UnifiedWalker<false, NESTED, true, 2> wk;
            wk.w = h.w;   // the pass-1 winner (a non-mesh object) and its rank
            wk.w.h.b0 = wk.w.h.b1 = wk.w.h.b2 = 0.0f;
            int stack_code[FW_STACK];
            float stack_te[FW_STACK];
            if (wk.init(S, h.o, h.d, stack_code, stack_te, nullptr)) {
                while (wk.step(S, key, nullptr)) {
                }
            }
            h.w = wk.w;
            mine = classify_winner(S, h.w, material);
        }
        enqueue_hit<MAT_NUM_QUEUES>(ps, s_fill, blockIdx.x * ps.seg_cap, mine, h.o, h.d, h.path, h.w, material);
    }
    seg_close<MAT_NUM_QUEUES>(s_fill, ps, counter_row(ps, bounce, 0), blockIdx.x);
}

template <bool NESTED, bool MESHES>
__global__ void __launch_bounds__(FW_BLOCK, FW_EXTEND_MIN_BLOCKS) extend_bvh_simple_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed,
                                                                                     uint32_t bounce) {
    FW_EXTEND_PROLOGUE(MAT_NUM_QUEUES)
        if (valid) {
            RngKey key{seed, 0u, 0u, bounce};
            batch_path(b, path, key.pixel, key.sample);
            trace_unified<false, NESTED, MESHES>(S, o, d, key, w, nullptr);
        }
    FW_EXTEND_EPILOGUE(MAT_NUM_QUEUES)
}

// Linear-scan scenes (Renderer.use_bvh == false, scene.rs:137-149): every ray tests every object in scene
// order, so there is no traversal-length divergence to balance.
// NESTED: the scene contains a TriangleMesh (its own BVH is walked inside the object test).
// This object-loop form serves scenes whose LinProgram does not fit kernel-parameter space.
template <bool NESTED>
__global__ void __launch_bounds__(FW_BLOCK) extend_linear_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed, uint32_t bounce) {
    FW_EXTEND_PROLOGUE(MAT_NUM_QUEUES)
        if (valid) {
            RngKey key{seed, 0u, 0u, bounce};
            batch_path(b, path, key.pixel, key.sample);
            trace_linear_scan<false, NESTED>(S, o, d, key, w, nullptr);
        }
    FW_EXTEND_EPILOGUE(MAT_NUM_QUEUES)
}

// The same query driven by the scene's LinProgram in kernel-parameter space (intersect.cuh trace_linear_prog):
// no per-lane loads of scene records, uniform item dispatch.  Every lane of a warp runs the program (lanes past
// the end of the segment trace a dummy ray and drop the result) so that the PRETEST vote sees the whole warp.
template <bool GENERIC, bool NESTED, bool PRETEST, bool SHDIV>
__global__ void __launch_bounds__(FW_BLOCK) extend_linear_prog_kernel(const __grid_constant__ LinProgram P, DeviceScene S, PathState ps,
                                                                      Batch b, uint2 seed, uint32_t bounce) {
    FW_EXTEND_PROLOGUE(MAT_NUM_QUEUES)
        RngKey key{seed, 0u, 0u, bounce};
        if (GENERIC) batch_path(b, path, key.pixel, key.sample);
        trace_linear_prog<false, GENERIC, NESTED, PRETEST, SHDIV>(P, S, o, d, key, w, nullptr);
    FW_EXTEND_EPILOGUE(MAT_NUM_QUEUES)
}

// render.rs:23 evaluated without recursion: colour = a0 * (a1 * (... (a_{k-1} * terminal))) with the same
// right-nested association as `emit + attenuation * color(...)`; emit is zero at every scattering vertex
// (material.rs:13-15), so the chain of attenuations is all that is needed.
FW_DEV float3 fold_radiance(const PathState& ps, uint32_t path, uint32_t bounce, float3 terminal) {
    // all loads first (independent, one HBM round trip), then the multiplications in the reference's order
    float4 a[FW_MAX_DEPTH];
#pragma unroll
    for (int k = 0; k < FW_MAX_DEPTH; ++k)
        if (k < (int)bounce) a[k] = ld_stream(&ps.atten[(size_t)k * ps.cap + path]);
    float3 x = terminal;
#pragma unroll
    for (int k = FW_MAX_DEPTH - 1; k >= 0; --k)
        if (k < (int)bounce) x = f3(a[k].x, a[k].y, a[k].z) * x;
    return x;
}

// render.rs:31 — environment lookup for rays that left the scene
__global__ void __launch_bounds__(FW_BLOCK) miss_kernel(DeviceScene S, PathState ps, uint32_t bounce) {
    const uint32_t total = counter_row(ps, bounce, MAT_MISS)[blockIdx.x];
    const float4* qd = ps.hq[MAT_MISS].d + (size_t)blockIdx.x * ps.seg_cap;
    for (uint32_t i = threadIdx.x; i < total; i += FW_BLOCK) {
        float4 rec = ld_stream(&qd[i]);
        uint32_t path = __float_as_uint(rec.w);
        float3 env = environment_sample(S.env, f3(rec));
        float3 c = fold_radiance(ps, path, bounce, env);
        st_stream(&ps.radiance[path], make_float4(c.x, c.y, c.z, 0.0f));
    }
}

// render.rs:20,25-28 with material.rs:174-180 — emissive surfaces end the path with their texture value
__global__ void __launch_bounds__(FW_BLOCK) shade_emissive_kernel(DeviceScene S, PathState ps, uint32_t bounce) {
    const uint32_t total = counter_row(ps, bounce, MAT_EMISSIVE)[blockIdx.x];
    const uint32_t base = blockIdx.x * ps.seg_cap;
    for (uint32_t i = threadIdx.x; i < total; i += FW_BLOCK) {
        HitIn h = get_hit<MAT_EMISSIVE>(ps, base + i);
        HitRecord rec;
        finalize_hit(S, h.w, h.o, h.d, rec, true);
        int tex = __ldg(&S.mats[h.material].tex);
        float3 emit = texture_sample(S, tex, rec.uv, rec.point);
        float3 c = fold_radiance(ps, h.path, bounce, emit);
        st_stream(&ps.radiance[h.path], make_float4(c.x, c.y, c.z, 0.0f));
    }
}

// Scattering materials: the next ray goes into the next extend queue, this vertex's attenuation into the chain.
// Not launched for bounce == FW_MAX_DEPTH (render.rs:21: no scatter at depth 10; emit is zero).
// The material kernels of one bounce run back to back and append to the same regions of the next extend queue.
template <int MAT>
__global__ void __launch_bounds__(FW_BLOCK, FW_SHADE_MIN_BLOCKS) shade_scatter_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed, uint32_t bounce) {
    __shared__ uint32_t s_fill[1];
    const uint32_t seg = blockIdx.x;
    const uint32_t total = counter_row(ps, bounce, MAT)[seg];
    if (total == 0) return;  // block-uniform
    uint32_t* row_out = counter_row(ps, bounce + 1, FW_Q_EXTEND);
    const uint32_t base = seg * ps.seg_cap;
    float4* __restrict__ xo = ps.xo[(bounce + 1) & 1];
    float4* __restrict__ xd = ps.xd[(bounce + 1) & 1];
    seg_open<1>(s_fill, ps, row_out, seg);
    for (uint32_t e0 = 0; e0 < total; e0 += FW_BLOCK) {
        uint32_t i = e0 + threadIdx.x;
        int mine = -1;
        uint32_t path = 0;
        ScatterOut out;
        out.scattered = false;
        out.origin = out.dir = out.attenuation = f3(0.0f, 0.0f, 0.0f);
        if (i + FW_BLOCK < total) {
            prefetch_l2(&ps.hq[MAT].o[base + i + FW_BLOCK]); prefetch_l2(&ps.hq[MAT].d[base + i + FW_BLOCK]);
            prefetch_l2(&ps.hq[MAT].w[base + i + FW_BLOCK]);
        }
        if (i < total) {
            HitIn h = get_hit<MAT>(ps, base + i);
            path = h.path;
            const float4* mq = reinterpret_cast<const float4*>(&S.mats[h.material]);
            float4 m0 = __ldg(mq), m1 = __ldg(mq + 1);  // (kind, tex, param, needs_uv), (albedo, -)
            HitRecord rec;
            finalize_hit(S, h.w, h.o, h.d, rec, __float_as_int(m0.w) != 0);
            float3 point = rec.point, normal = rec.normal;
            uint32_t pixel, sample;
            batch_path(b, path, pixel, sample);
            RngKey key{seed, pixel, sample, bounce};
            PhiloxStream rng(key, STREAM_SCATTER);
            if (MAT == MAT_LAMBERTIAN) {
                scatter_lambertian(S, __float_as_int(m0.y), point, normal, rec.uv, rng, out);
            } else if (MAT == MAT_METAL) {
                scatter_metal(f3(m1), m0.z, h.d, point, normal, rng, out);
            } else if (MAT == MAT_DIELECTRIC) {
                scatter_dielectric(m0.z, h.d, point, normal, rng, out);
            } else {
                scatter_isotropic(S, __float_as_int(m0.y), point, rec.uv, rng, out);
            }
            if (out.scattered) {
                st_stream(&ps.atten[(size_t)bounce * ps.cap + path],
                          make_float4(out.attenuation.x, out.attenuation.y, out.attenuation.z, 0.0f));
                mine = 0;
            }
            // absorbed (metal below the surface): radiance stays 0 (render.rs:25)
        }
        uint32_t slot = seg_reserve<1>(s_fill, base, mine);
        if (mine == 0) {
            st_stream(&xo[slot], make_float4(out.origin.x, out.origin.y, out.origin.z, __uint_as_float(path)));
            st_stream(&xd[slot], make_float4(out.dir.x, out.dir.y, out.dir.z, 0.0f));
        }
    }
    seg_close<1>(s_fill, ps, row_out, seg);
}

// Ray statistics without a host round trip per batch: rays traced = every entry of every bounce's extend queue.
__global__ void __launch_bounds__(256) tally_kernel(PathState ps, unsigned long long* total_rays) {
    __shared__ unsigned long long s_part[256];
    unsigned long long r = 0;
    for (int bn = 0; bn <= FW_MAX_DEPTH; ++bn) {   // row b = rays traced at bounce b (row 0 = the primary rays)
        const uint32_t* row = counter_row(ps, bn, FW_Q_EXTEND);
        for (uint32_t i = threadIdx.x; i < ps.nseg; i += 256) r += row[i];
    }
    s_part[threadIdx.x] = r;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) s_part[threadIdx.x] += s_part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_rays += s_part[0];
}

// render.rs:177-182 `total_color += color(...)`: samples of a pixel are added in sample order, so the fp32
// sum is independent of queue order and identical for any batch split along the sample axis.
__global__ void __launch_bounds__(256) accumulate_kernel(float* __restrict__ sum, PathState ps, Batch b) {
    for (uint32_t pl = blockIdx.x * blockDim.x + threadIdx.x; pl < b.npix; pl += gridDim.x * blockDim.x) {
        size_t pix = (size_t)b.pix0 + pl;
        float r = sum[3 * pix], g = sum[3 * pix + 1], bl = sum[3 * pix + 2];
        for (uint32_t s = 0; s < b.ns; ++s) {
            float4 c = ld_stream(&ps.radiance[(size_t)s * b.npix + pl]);
            r += c.x; g += c.y; bl += c.z;
        }
        sum[3 * pix] = r; sum[3 * pix + 1] = g; sum[3 * pix + 2] = bl;
    }
}

// render.rs:184-189 + util.rs:14-23: mean, powf(1/gamma), clamp, *255.99 -> saturating u8 (NaN -> 0)
FW_DEV unsigned char quantise(float mean, float inv_gamma) {
    float x = powf(mean, inv_gamma);
    if (x < 0.0f) x = 0.0f;
    if (x > 1.0f) x = 1.0f;
    float y = x * 255.99f;
    if (!(y == y)) return 0;
    return (unsigned char)fminf(fmaxf(truncf(y), 0.0f), 255.0f);
}
__global__ void __launch_bounds__(256) resolve_kernel(const float* __restrict__ sum, uint32_t npix, float samples,
                                                      float gamma, unsigned char* __restrict__ rgb) {
    float inv_gamma = 1.0f / gamma;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        rgb[3 * i] = quantise(sum[3 * i] / samples, inv_gamma);
        rgb[3 * i + 1] = quantise(sum[3 * i + 1] / samples, inv_gamma);
        rgb[3 * i + 2] = quantise(sum[3 * i + 2] / samples, inv_gamma);
    }
}

// ---- probe kernels (the parity gates of BASELINE.json) -------------------------------------------------

__global__ void primary_rays_probe(CameraRec cam, uint32_t width, uint32_t height, uint32_t sample, uint2 seed,
                                   uint32_t pix_begin, uint32_t n, float* origins, float* dirs) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float3 o, d;
        primary_ray(cam, width, height, pix_begin + i, sample, seed, o, d);
        origins[3 * i] = o.x; origins[3 * i + 1] = o.y; origins[3 * i + 2] = o.z;
        dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z;
    }
}

struct FirstHitOut {
    int *obj, *prim, *material;
    float *t, *point, *normal, *uv;
    unsigned long long* counters;  // [2] node tests, prim tests
};
template <bool USE_BVH>
__global__ void first_hit_probe(DeviceScene S, uint2 seed, uint32_t n, const float* origins, const float* dirs,
                                const uint32_t* pixel, const uint32_t* sample, const uint32_t* bounce,
                                FirstHitOut out) {
    Counters cnt{0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float3 o = f3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        float3 d = f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        RngKey key{seed, pixel ? pixel[i] : i, sample ? sample[i] : 0u, bounce ? bounce[i] : 0u};
        HitRecord rec;
        if (scene_closest_hit<USE_BVH, true>(S, o, d, key, rec, &cnt)) {
            out.obj[i] = rec.obj; out.prim[i] = rec.prim; out.material[i] = rec.material; out.t[i] = rec.t;
            out.point[3 * i] = rec.point.x; out.point[3 * i + 1] = rec.point.y; out.point[3 * i + 2] = rec.point.z;
            out.normal[3 * i] = rec.normal.x; out.normal[3 * i + 1] = rec.normal.y; out.normal[3 * i + 2] = rec.normal.z;
            out.uv[2 * i] = rec.uv.x; out.uv[2 * i + 1] = rec.uv.y;
        } else {
            out.obj[i] = -1; out.prim[i] = 0; out.material[i] = -1; out.t[i] = 0.0f;
            out.point[3 * i] = out.point[3 * i + 1] = out.point[3 * i + 2] = 0.0f;
            out.normal[3 * i] = out.normal[3 * i + 1] = out.normal[3 * i + 2] = 0.0f;
            out.uv[2 * i] = out.uv[2 * i + 1] = 0.0f;
        }
    }
    atomicAdd(&out.counters[0], cnt.node_tests);
    atomicAdd(&out.counters[1], cnt.prim_tests);
}

// first-hit probe through the LinProgram path (what linear-scan renders execute)
template <bool GENERIC>
__global__ void first_hit_prog_probe(const __grid_constant__ LinProgram P, DeviceScene S, uint2 seed, uint32_t n, const float* origins,
                                     const float* dirs, const uint32_t* pixel, const uint32_t* sample, const uint32_t* bounce,
                                     FirstHitOut out) {
    Counters cnt{0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float3 o = f3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        float3 d = f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        RngKey key{seed, pixel ? pixel[i] : i, sample ? sample[i] : 0u, bounce ? bounce[i] : 0u};
        Winner w;
        trace_linear_prog<true, GENERIC, true, false, !GENERIC>(P, S, o, d, key, w, &cnt);
        if (w.found) {
            HitRecord rec;
            finalize_hit(S, w, o, d, rec);
            out.obj[i] = rec.obj; out.prim[i] = rec.prim; out.material[i] = rec.material; out.t[i] = rec.t;
            out.point[3 * i] = rec.point.x; out.point[3 * i + 1] = rec.point.y; out.point[3 * i + 2] = rec.point.z;
            out.normal[3 * i] = rec.normal.x; out.normal[3 * i + 1] = rec.normal.y; out.normal[3 * i + 2] = rec.normal.z;
            out.uv[2 * i] = rec.uv.x; out.uv[2 * i + 1] = rec.uv.y;
        } else {
            out.obj[i] = -1; out.prim[i] = 0; out.material[i] = -1; out.t[i] = 0.0f;
            out.point[3 * i] = out.point[3 * i + 1] = out.point[3 * i + 2] = 0.0f;
            out.normal[3 * i] = out.normal[3 * i + 1] = out.normal[3 * i + 2] = 0.0f;
            out.uv[2 * i] = out.uv[2 * i + 1] = 0.0f;
        }
    }
    atomicAdd(&out.counters[0], cnt.node_tests);
    atomicAdd(&out.counters[1], cnt.prim_tests);
}

struct ScatterProbeIO {
    const int* material;
    const float *ray_o, *ray_d, *hit_t, *hit_point, *hit_normal, *hit_uv, *uniforms;
    uint32_t nu;
    float* emit;
    int* scattered;
    float *atten, *out_o, *out_d;
    int* consumed;
};
__global__ void scatter_step_probe(DeviceScene S, uint32_t n, ScatterProbeIO io) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int mat = io.material[i];
        const float4* mq = reinterpret_cast<const float4*>(&S.mats[mat]);
        float4 m0 = __ldg(mq), m1 = __ldg(mq + 1);
        int kind = __float_as_int(m0.x), tex = __float_as_int(m0.y);
        float3 in_d = f3(io.ray_d[3 * i], io.ray_d[3 * i + 1], io.ray_d[3 * i + 2]);
        float3 point = f3(io.hit_point[3 * i], io.hit_point[3 * i + 1], io.hit_point[3 * i + 2]);
        float3 normal = f3(io.hit_normal[3 * i], io.hit_normal[3 * i + 1], io.hit_normal[3 * i + 2]);
        float2 uv = make_float2(io.hit_uv[2 * i], io.hit_uv[2 * i + 1]);
        ArrayStream rng(io.uniforms + (size_t)i * io.nu, (int)io.nu);
        ScatterOut out;
        out.scattered = false;
        out.attenuation = out.origin = out.dir = f3(0.0f, 0.0f, 0.0f);
        float3 emit = f3(0.0f, 0.0f, 0.0f);
        switch (kind) {
            case MAT_LAMBERTIAN: scatter_lambertian(S, tex, point, normal, uv, rng, out); break;
            case MAT_METAL: scatter_metal(f3(m1), m0.z, in_d, point, normal, rng, out); break;
            case MAT_DIELECTRIC: scatter_dielectric(m0.z, in_d, point, normal, rng, out); break;
            case MAT_EMISSIVE: emit = texture_sample(S, tex, uv, point); break;
            case MAT_ISOTROPIC: scatter_isotropic(S, tex, point, uv, rng, out); break;
        }
        if (!out.scattered) out.attenuation = out.origin = out.dir = f3(0.0f, 0.0f, 0.0f);
        io.emit[3 * i] = emit.x; io.emit[3 * i + 1] = emit.y; io.emit[3 * i + 2] = emit.z;
        io.scattered[i] = out.scattered ? 1 : 0;
        io.atten[3 * i] = out.attenuation.x; io.atten[3 * i + 1] = out.attenuation.y; io.atten[3 * i + 2] = out.attenuation.z;
        io.out_o[3 * i] = out.origin.x; io.out_o[3 * i + 1] = out.origin.y; io.out_o[3 * i + 2] = out.origin.z;
        io.out_d[3 * i] = out.dir.x; io.out_d[3 * i + 1] = out.dir.y; io.out_d[3 * i + 2] = out.dir.z;
        io.consumed[i] = rng.overrun ? -1 : rng.i;
    }
}

// Self-test of intersect.cuh div_by / shared_div against the hardware's IEEE `/`: pseudo-random and adversarial
// operands around and inside the fast domain.  violations[0] counts results that differ in any bit where the helper
// promises the exact quotient (|n| >= 2^-60 or outside the fast domain), violations[1] counts tiny-numerator cases
// where either value reaches the only threshold it is ever compared with (t_min = 0.001).
__global__ void shared_division_probe(uint64_t n_pairs, uint2 seed, unsigned long long* violations) {
    unsigned long long bad = 0, bad_tiny = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0x5d1fu, 7u), seed);
        // mantissas: random, or one of the adversarial patterns
        const uint32_t pat[8] = {0x000000u, 0x7fffffu, 0x000001u, 0x7ffffeu, 0x400000u, 0x3fffffu, 0x555555u, 0x2aaaaau};
        uint32_t mn = (r.z & 8u) ? pat[r.z & 7u] : (r.x & 0x7fffffu);
        uint32_t md = (r.z & 128u) ? pat[(r.z >> 4) & 7u] : (r.y & 0x7fffffu);
        int en = -70 + (int)((r.w & 0xffffu) % 133u);        // 2^-70 .. 2^62
        int ed = -42 + (int)((r.w >> 16) % 85u);              // 2^-42 .. 2^42
        uint32_t sn = (r.x >> 31) << 31, sd = (r.y >> 31) << 31;
        float n = __uint_as_float(sn | ((uint32_t)(en + 127) << 23) | mn);
        float d = __uint_as_float(sd | ((uint32_t)(ed + 127) << 23) | md);
        if ((r.z & 0xff00u) == 0x1100u) n = __uint_as_float(sn);   // exact zero numerators
        float want = n / d;
        float got = div_by<true>(n, d, shared_div(d));
        bool in_domain = fabsf(d) >= 9.094947017729282e-13f && fabsf(d) <= 1.099511627776e12f && fabsf(n) <= 1.152921504606846976e18f;
        bool tiny = in_domain && fabsf(n) < 8.673617379884035e-19f;   // 2^-60
        if (tiny) {
            if (!(fabsf(want) < 0.001f) || !(fabsf(got) < 0.001f)) ++bad_tiny;
        } else if (__float_as_uint(want) != __float_as_uint(got)) {
            ++bad;
        }
    }
    if (bad) atomicAdd(&violations[0], bad);
    if (bad_tiny) atomicAdd(&violations[1], bad_tiny);
}

__global__ void env_sample_probe(DeviceScene S, uint32_t n, const float* dirs, float* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float3 c = environment_sample(S.env, f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
__global__ void texture_sample_probe(DeviceScene S, int tex, uint32_t n, const float* uv, const float* point, float* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float3 c = texture_sample(S, tex, make_float2(uv[2 * i], uv[2 * i + 1]),
                                  f3(point[3 * i], point[3 * i + 1], point[3 * i + 2]));
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}

}  // namespace fw
