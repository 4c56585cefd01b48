// Collect / test BVH traversal kernels (see walk.cuh for the design): walk_top_kernel, walk_mesh_kernel, classify_kernel.
#include "launch.h"
#include "walk.cuh"

namespace fw {

// Winner of a ray from its final key (found / t / object / rank / primitive).
FW_DEV void winner_from_key(const DeviceScene& S, unsigned long long k, int prim_bits, Winner& w) {
    w.found = k != FW_KEY_NONE;
    w.t = 0.0f; w.obj = -1; w.rank = -1;
    w.h.t = 0.0f; w.h.prim = 0; w.h.b0 = w.h.b1 = w.h.b2 = 0.0f;
    if (w.found) {
        w.t = w.h.t = key_t(k);
        w.rank = key_rank(k, prim_bits);
        w.h.prim = key_prim(k, prim_bits);
        w.obj = __ldg(&S.leaf_meta[w.rank]).w;
    }
}

// ---- top level -------------------------------------------------------------------------------------------------
// MESHES: the scene has TriangleMesh render objects: rays are not classified here; their keys go to aux.tkey and the
//         (ray, mesh) entries to aux.entries (walk_mesh_kernel + classify_kernel finish the bounce).
// NESTED: some ConstantMedium wraps a TriangleMesh (its tree is walked inside the shape test).
template <bool MESHES, bool NESTED>
__global__ void __launch_bounds__(FW_BLOCK, FW_WALK_MIN_BLOCKS) walk_top_kernel(DeviceScene S, PathState ps, Batch b, uint2 seed,
                                                                                uint32_t bounce, WalkAux aux) {
    __shared__ uint32_t s_fill[FW_NUM_QUEUES];
    __shared__ uint32_t s_cursor;
    __shared__ WalkWarp s_w[FW_WALK_WARPS];
    const uint32_t seg = blockIdx.x;
    const uint32_t in_count = counter_row(ps, bounce, FW_Q_EXTEND)[seg];
    if (in_count == 0) return;   // block-uniform
    seg_open<FW_NUM_QUEUES>(s_fill, ps, counter_row(ps, bounce, 0), seg);
    if (threadIdx.x == 0) s_cursor = 0;
    __syncthreads();
    WalkWarp& W = s_w[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const size_t seg_base = (size_t)seg * ps.seg_cap;
    const float4* __restrict__ qo = ps.xo[bounce & 1] + seg_base;
    const float4* __restrict__ qd = ps.xd[bounce & 1] + seg_base;
    const float tmin = 0.001f, tmax = 2e9f;   // render.rs:19
    const int pb = aux.prim_bits;
    const bool no_cull = S.has_unbounded != 0;   // a Disk's box does not bound it (disk.rs:85-90): never cull by distance

    bool has_ray = false, active = false, nan_ray = false;
    float3 o = f3(0.0f, 0.0f, 0.0f), inv = f3(1.0f, 1.0f, 1.0f);
    int node = -1, sp = 0;
    uint32_t slot_in = 0;
    int stk_code[FW_WALK_STACK];
    float stk_te[FW_WALK_STACK];
    int npairs = 0;
    bool input_left = true;

    // one pair per lane: the leaf's items against the owning lane's ray (bvh.rs:119-133 over render objects)
    auto test_pairs = [&](int n) {
        __syncwarp();
        if ((int)lane < n) {
            const uint32_t pr = W.pairs[npairs - n + lane];
            const unsigned src = pr & 31u;
            const int packed = (int)(pr >> 5);
            const int first = packed >> 1, count = (packed & 1) + 1;
            const float3 ro = f3(W.ox[src], W.oy[src], W.oz[src]), rd = f3(W.dx[src], W.dy[src], W.dz[src]);
            RngKey key{seed, 0u, 0u, bounce};
            if (S.has_medium) batch_path(b, W.a0[src], key.pixel, key.sample);
            for (int k = 0; k < count; ++k) {
                const int rank = first + k;
                const float4 posr = __ldg(&S.leaf_posr[rank]);
                const int4 meta = __ldg(&S.leaf_meta[rank]);
                const float bnd = no_cull ? FW_FLT_MAX : key_bound(W.key[src]);
                if (MESHES && (meta.x & OBJ_KIND_MASK) == SH_MESH) {
                    // scene.rs:242-253 + the mesh root's box test (bvh.rs:117); the walk itself is walk_mesh_kernel's
                    float3 oo = ro - f3(posr), od = rd;
                    if (meta.x & OBJ_ROTATED) {
                        const float4* m = &S.obj_irot[3 * meta.w];
                        float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
                        oo = mat_mul(r0, r1, r2, oo);
                        od = mat_mul(r0, r1, r2, rd);
                    }
                    const float4* q = reinterpret_cast<const float4*>(&S.shapes[meta.z]);
                    const float4* mr = reinterpret_cast<const float4*>(&S.meshes[as_int(__ldg(q).z)]);
                    const float3 oinv = f3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
                    float te;
                    if (slab_test(__ldg(mr + 2), __ldg(mr + 3), oo, oinv, tmin, tmax, te) && !(te > bnd)) {
                        const uint32_t e = atomicAdd(&s_fill[FW_Q_MESH], 1u);
                        aux.entries[(size_t)seg * aux.ent_cap + e] = make_uint2(W.a1[src], (uint32_t)rank);
                    }
                    continue;
                }
                ObjHit h;
                if (object_test_loaded<false, NESTED>(S, meta.w, posr, meta, ro, rd, tmin, tmax, bnd, key, h, nullptr))
                    atomicMin(&W.key[src], pack_key(h.t, rank, h.prim, pb));
            }
        }
        npairs -= n;
        __syncwarp();
    };

    for (;;) {
        unsigned act = __ballot_sync(0xffffffffu, active);
        if ((input_left && 32 - __popc(act) >= FW_WALK_REFILL_IDLE) || act == 0u) {
            // ---- flush the pair buffer: the rays about to retire may still own pairs
            while (npairs > 0) test_pairs(npairs < 32 ? npairs : 32);
            // ---- retire the rays that have no node left
            const bool retiring = has_ray && !active;
            if (MESHES) {
                if (retiring) aux.tkey[seg_base + slot_in] = W.key[lane];
            } else {
                Winner w;
                w.found = false; w.t = 0.0f; w.obj = -1; w.rank = -1; w.h.t = 0.0f; w.h.prim = 0; w.h.b0 = w.h.b1 = w.h.b2 = 0.0f;
                int mine = -1, material = -1;
                float3 ro = f3(0.0f, 0.0f, 0.0f), rd = ro;
                uint32_t path = 0;
                if (retiring) {
                    ro = f3(W.ox[lane], W.oy[lane], W.oz[lane]); rd = f3(W.dx[lane], W.dy[lane], W.dz[lane]);
                    path = W.a0[lane];
                    if (nan_ray) nan_direction_winner(S.nan_bvh_obj, S.nan_bvh_prim, w);
                    else winner_from_key(S, W.key[lane], pb, w);
                    mine = classify_winner(S, w, material);
                }
                enqueue_hit<MAT_NUM_QUEUES>(ps, s_fill, blockIdx.x * ps.seg_cap, mine, ro, rd, path, w, material);
            }
            if (retiring) has_ray = false;
            // ---- fetch new rays for the free lanes
            if (input_left) {
                const unsigned want = __ballot_sync(0xffffffffu, !has_ray);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_cursor, (uint32_t)__popc(want));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (!has_ray) {
                    const uint32_t e = base + __popc(want & lt);
                    if (e < in_count) {
                        const float4 ro = ld_stream(&qo[e]), rd = ld_stream(&qd[e]);
                        o = f3(ro);
                        const float3 d = f3(rd);
                        W.ox[lane] = o.x; W.oy[lane] = o.y; W.oz[lane] = o.z;
                        W.dx[lane] = d.x; W.dy[lane] = d.y; W.dz[lane] = d.z;
                        W.a0[lane] = __float_as_uint(ro.w);   // path
                        W.a1[lane] = e;
                        W.key[lane] = FW_KEY_NONE;
                        slot_in = e;
                        has_ray = true;
                        nan_ray = nan_direction(d);
                        sp = 0;
                        node = -1;
                        if (!nan_ray) {
                            inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                            float te;
                            if (slab_test(S.top_lo, S.top_hi, o, inv, tmin, tmax, te)) {   // bvh.rs:117 at the root
                                node = as_int(S.top_lo.w);
                                active = true;
                            }
                        }
                    }
                }
                if (base + (uint32_t)__popc(want) >= in_count) input_left = false;
            }
            act = __ballot_sync(0xffffffffu, active);
            if (act == 0u) {
                if (!input_left && __ballot_sync(0xffffffffu, has_ray) == 0u) break;
                continue;   // every fetched ray missed the root box: retire them and fetch again
            }
        }
        // ---- walk: one wide-node visit per lane that has a node
        bool l0 = false, l1 = false, l2 = false, l3 = false;
        int4 cc = make_int4(0, 0, 0, 0);
        if (active) {
            const float bnd = no_cull ? FW_FLT_MAX : key_bound(W.key[lane]);
            if (node < 0) {
                for (;;) {
                    if (sp == 0) { active = false; break; }
                    --sp;
                    if (!(stk_te[sp] > bnd)) { node = stk_code[sp]; break; }
                }
            }
            if (active) walk_visit(S.nodes, node, o, inv, bnd, stk_code, stk_te, sp, l0, l1, l2, l3, cc);
        }
        npairs = walk_emit(W, npairs, l0, l1, l2, l3, cc);
        while (npairs >= 32) test_pairs(32);
    }
    seg_close<FW_NUM_QUEUES>(s_fill, ps, counter_row(ps, bounce, 0), seg);
}

// ---- mesh level ------------------------------------------------------------------------------------------------
// One work item = one (ray, mesh) entry: the mesh's tree is walked in the mesh's object space (scene.rs:242-253),
// triangles are tested in the test phase (mesh.rs:140-219), and the entry's best (t, rank, slot) is merged into the
// ray's key in global memory.
__global__ void __launch_bounds__(FW_BLOCK, FW_WALK_MIN_BLOCKS) walk_mesh_kernel(DeviceScene S, PathState ps, uint32_t bounce,
                                                                                 WalkAux aux) {
    __shared__ uint32_t s_cursor;
    __shared__ WalkWarp s_w[FW_WALK_WARPS];
    const uint32_t seg = blockIdx.x;
    const uint32_t in_count = counter_row(ps, bounce, FW_Q_MESH)[seg];
    if (in_count == 0) return;   // block-uniform
    if (threadIdx.x == 0) s_cursor = 0;
    __syncthreads();
    WalkWarp& W = s_w[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const size_t seg_base = (size_t)seg * ps.seg_cap;
    const float4* __restrict__ qo = ps.xo[bounce & 1] + seg_base;
    const float4* __restrict__ qd = ps.xd[bounce & 1] + seg_base;
    const uint2* __restrict__ ents = aux.entries + (size_t)seg * aux.ent_cap;
    unsigned long long* __restrict__ tkey = aux.tkey + seg_base;
    const float tmin = 0.001f, tmax = 2e9f;
    const int pb = aux.prim_bits;

    bool has_ray = false, active = false;
    float3 o = f3(0.0f, 0.0f, 0.0f), inv = f3(1.0f, 1.0f, 1.0f);
    int node = -1, sp = 0;
    uint32_t slot_in = 0;
    int stk_code[FW_WALK_STACK];
    float stk_te[FW_WALK_STACK];
    int npairs = 0;
    bool input_left = true;

    // one pair per lane: the leaf's 1-2 triangles against the owning entry's ray (bvh.rs:119-133 over Triangle items)
    auto test_pairs = [&](int n) {
        __syncwarp();
        if ((int)lane < n) {
            const uint32_t pr = W.pairs[npairs - n + lane];
            const unsigned src = pr & 31u;
            const int packed = (int)(pr >> 5);
            const int first = packed >> 1, count = (packed & 1) + 1;
            const float3 ro = f3(W.ox[src], W.oy[src], W.oz[src]);
            TriSetup su;
            su.kz = W.su_k[src]; su.sx = W.su_x[src]; su.sy = W.su_y[src]; su.sz = W.su_z[src];
            const int tri_first = (int)W.a0[src], rank = (int)W.a1[src];
            for (int k = 0; k < count; ++k) {
                const int slot = first + k;
                const float4* v = &S.tri_verts[3 * (tri_first + slot)];
                const float4 q0 = __ldg(v), q1 = __ldg(v + 1), q2 = __ldg(v + 2);
                float t, c0, c1, c2;
                if (triangle_test(f3(q0), f3(q1), f3(q2), ro, su, tmin, tmax, t, c0, c1, c2))
                    atomicMin(&W.key[src], pack_key(t, rank, slot, pb));
            }
        }
        npairs -= n;
        __syncwarp();
    };

    for (;;) {
        unsigned act = __ballot_sync(0xffffffffu, active);
        if ((input_left && 32 - __popc(act) >= FW_WALK_REFILL_IDLE) || act == 0u) {
            while (npairs > 0) test_pairs(npairs < 32 ? npairs : 32);
            if (has_ray && !active) {   // retire: merge this entry's best into the ray's key
                const unsigned long long k = W.key[lane];
                if (k != FW_KEY_NONE) atomicMin(&tkey[slot_in], k);
                has_ray = false;
            }
            if (input_left) {
                const unsigned want = __ballot_sync(0xffffffffu, !has_ray);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_cursor, (uint32_t)__popc(want));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (!has_ray) {
                    const uint32_t e = base + __popc(want & lt);
                    if (e < in_count) {
                        const uint2 en = ents[e];
                        slot_in = en.x;
                        const int rank = (int)en.y;
                        const float4 ro = __ldg(&qo[slot_in]), rd = __ldg(&qd[slot_in]);
                        const float4 posr = __ldg(&S.leaf_posr[rank]);
                        const int4 meta = __ldg(&S.leaf_meta[rank]);
                        float3 oo = f3(ro) - f3(posr), od = f3(rd);
                        if (meta.x & OBJ_ROTATED) {   // scene.rs:242-249
                            const float4* m = &S.obj_irot[3 * meta.w];
                            float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
                            od = mat_mul(r0, r1, r2, f3(rd));
                            oo = mat_mul(r0, r1, r2, oo);
                        }
                        const float4* q = reinterpret_cast<const float4*>(&S.shapes[meta.z]);
                        const float4* mr = reinterpret_cast<const float4*>(&S.meshes[as_int(__ldg(q).z)]);
                        const int4 m0 = __ldg(reinterpret_cast<const int4*>(mr));
                        o = oo;
                        inv = f3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
                        const TriSetup su = tri_setup(od);
                        W.ox[lane] = oo.x; W.oy[lane] = oo.y; W.oz[lane] = oo.z;
                        W.su_k[lane] = su.kz; W.su_x[lane] = su.sx; W.su_y[lane] = su.sy; W.su_z[lane] = su.sz;
                        W.a0[lane] = (uint32_t)m0.y;   // first triangle slot
                        W.a1[lane] = (uint32_t)rank;
                        // the ray's best so far (top-level winner, or an earlier entry's triangle): only a culling bound here
                        W.key[lane] = *reinterpret_cast<volatile unsigned long long*>(&tkey[slot_in]);
                        has_ray = true;
                        active = true;
                        sp = 0;
                        node = m0.x;   // the mesh's root (a wide node: checked at flatten); its box was tested at top level
                    }
                }
                if (base + (uint32_t)__popc(want) >= in_count) input_left = false;
            }
            act = __ballot_sync(0xffffffffu, active);
            if (act == 0u) {
                if (!input_left && __ballot_sync(0xffffffffu, has_ray) == 0u) break;
                continue;
            }
        }
        bool l0 = false, l1 = false, l2 = false, l3 = false;
        int4 cc = make_int4(0, 0, 0, 0);
        if (active) {
            const float bnd = key_bound(W.key[lane]);
            if (node < 0) {
                for (;;) {
                    if (sp == 0) { active = false; break; }
                    --sp;
                    if (!(stk_te[sp] > bnd)) { node = stk_code[sp]; break; }
                }
            }
            if (active) walk_visit(S.nodes, node, o, inv, bnd, stk_code, stk_te, sp, l0, l1, l2, l3, cc);
        }
        npairs = walk_emit(W, npairs, l0, l1, l2, l3, cc);
        while (npairs >= 32) test_pairs(32);
    }
}

// ---- classification (mesh scenes) ---------------------------------------------------------------------------------
// Every ray of the bounce: final key -> winner -> its material's shade queue (or the miss queue).
__global__ void __launch_bounds__(FW_BLOCK) classify_kernel(DeviceScene S, PathState ps, uint32_t bounce, WalkAux aux) {
    __shared__ uint32_t s_fill[FW_NUM_QUEUES];
    const uint32_t seg = blockIdx.x;
    const uint32_t in_count = counter_row(ps, bounce, FW_Q_EXTEND)[seg];
    if (in_count == 0) return;
    seg_open<MAT_NUM_QUEUES>(s_fill, ps, counter_row(ps, bounce, 0), seg);
    const size_t seg_base = (size_t)seg * ps.seg_cap;
    for (uint32_t e0 = 0; e0 < in_count; e0 += FW_BLOCK) {
        const uint32_t e = e0 + threadIdx.x;
        Winner w;
        w.found = false; w.t = 0.0f; w.obj = -1; w.rank = -1; w.h.t = 0.0f; w.h.prim = 0; w.h.b0 = w.h.b1 = w.h.b2 = 0.0f;
        int mine = -1, material = -1;
        float3 o = f3(0.0f, 0.0f, 0.0f), d = o;
        uint32_t path = 0;
        if (e < in_count) {
            const float4 ro = ld_stream(&ps.xo[bounce & 1][seg_base + e]), rd = ld_stream(&ps.xd[bounce & 1][seg_base + e]);
            o = f3(ro); d = f3(rd); path = __float_as_uint(ro.w);
            if (nan_direction(d)) nan_direction_winner(S.nan_bvh_obj, S.nan_bvh_prim, w);
            else winner_from_key(S, aux.tkey[seg_base + e], aux.prim_bits, w);
            mine = classify_winner(S, w, material);
        }
        enqueue_hit<MAT_NUM_QUEUES>(ps, s_fill, blockIdx.x * ps.seg_cap, mine, o, d, path, w, material);
    }
    seg_close<MAT_NUM_QUEUES>(s_fill, ps, counter_row(ps, bounce, 0), seg);
}

// ---- launcher ----------------------------------------------------------------------------------------------------
int launch_extend_walk(const ExtendPlan& plan, const DeviceScene& S, const PathState& ps, const Batch& b, uint2 seed,
                       uint32_t bounce, const WalkAuxHost& ax, cudaStream_t st) {
    WalkAux aux;
    aux.tkey = ax.tkey; aux.entries = reinterpret_cast<uint2*>(ax.entries); aux.ent_cap = ax.ent_cap; aux.prim_bits = ax.prim_bits;
    const unsigned G = ps.nseg;
    if (plan.has_top_mesh) {
        if (plan.has_medium_mesh) walk_top_kernel<true, true><<<G, FW_BLOCK, 0, st>>>(S, ps, b, seed, bounce, aux);
        else walk_top_kernel<true, false><<<G, FW_BLOCK, 0, st>>>(S, ps, b, seed, bounce, aux);
        walk_mesh_kernel<<<G, FW_BLOCK, 0, st>>>(S, ps, bounce, aux);
        classify_kernel<<<G, FW_BLOCK, 0, st>>>(S, ps, bounce, aux);
        return 3;
    }
    if (plan.has_medium_mesh) walk_top_kernel<false, true><<<G, FW_BLOCK, 0, st>>>(S, ps, b, seed, bounce, aux);
    else walk_top_kernel<false, false><<<G, FW_BLOCK, 0, st>>>(S, ps, b, seed, bounce, aux);
    return 1;
}

}  // namespace fw
