"""Thin host-side wrapper over the C ABI: scene handle, asset hand-over, render and probe calls.

Mirrors the call sequence of the reference's front ends (src/main.rs:22-47, examples/*.rs):
load/build a `Scene` -> `Renderer::render(scene)` -> pixels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _native as N
from .assets import load_asset


class NativeScene:
    """A committed fw_scene on one GPU."""

    def __init__(self, yaml_text: str, device: int = 0, asset_dir: Optional[str] = None,
                 assets: Optional[dict] = None, commit: bool = True):
        L = N.lib()
        self._h = C.c_void_p()
        data = yaml_text.encode("utf-8")
        N.check(L.fw_scene_from_yaml(data, len(data), C.byref(self._h)))
        self.device = device
        self.h2d_asset_bytes = 0
        for i in range(L.fw_scene_num_assets(self._h)):
            path = L.fw_scene_asset_path(self._h, i).decode()
            kind = L.fw_scene_asset_kind(self._h, i)
            arr = load_asset(path, "hdr" if kind == 1 else "image", asset_dir, assets)
            if kind == 1:
                arr = np.ascontiguousarray(arr, np.float32)
                N.check(L.fw_scene_set_hdr(self._h, i, arr.shape[1], arr.shape[0], N.ptr(arr)))
            else:
                arr = np.ascontiguousarray(arr, np.uint8)
                N.check(L.fw_scene_set_image(self._h, i, arr.shape[1], arr.shape[0], N.ptr(arr)))
            self.h2d_asset_bytes += arr.nbytes
        self.yaml_bytes = len(data)
        if commit:
            self.commit()

    @staticmethod
    def from_scene(scene, device: int = 0) -> "NativeScene":
        return NativeScene(scene.to_yaml(), device, scene.asset_dir, scene.assets)

    def commit(self):
        N.check(N.lib().fw_scene_commit(self._h, self.device))

    def build_host(self):
        """BVH build + flattening only (no CUDA): enough for the structural introspection calls."""
        N.check(N.lib().fw_scene_build_host(self._h))

    def close(self):
        if self._h:
            N.lib().fw_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ---------------------------------------------------------------------------------
    def num_objects(self):
        return N.lib().fw_scene_num_objects(self._h)

    def num_nodes(self):
        return N.lib().fw_scene_num_nodes(self._h)

    def device_bytes(self):
        return int(N.lib().fw_scene_device_bytes(self._h))

    def top_leaf_order(self):
        n = self.num_objects()
        out = np.zeros(n, np.int32)
        N.check(N.lib().fw_scene_top_leaf_order(self._h, N.ptr(out), n))
        return out

    def object_aabbs(self):
        n = self.num_objects()
        out = np.zeros((n, 6), np.float32)
        for i in range(n):
            N.check(N.lib().fw_scene_object_aabb(self._h, i, N.ptr(out[i])))
        return out

    def mesh_leaf_order(self, obj):
        n = N.check(N.lib().fw_scene_mesh_leaf_order(self._h, obj, None, 0))
        out = np.zeros(n, np.int32)
        N.check(N.lib().fw_scene_mesh_leaf_order(self._h, obj, N.ptr(out), n))
        return out

    def bvh_nodes(self):
        """(nodes (n, 8, 4) float32 — int fields are bit patterns —, top-level root code)."""
        root = C.c_int()
        n = N.check(N.lib().fw_scene_bvh_nodes(self._h, None, 0, C.byref(root)))
        out = np.zeros((max(n, 1), 8, 4), np.float32)
        N.check(N.lib().fw_scene_bvh_nodes(self._h, N.ptr(out), n, C.byref(root)))
        return out[:n], root.value

    def linear_program(self):
        """The scene's linear-scan program as an (n_words, 4) float32 array (fw_types.h LinItem encoding)."""
        n = N.check(N.lib().fw_scene_linear_program(self._h, None, 0))
        out = np.zeros((n, 4), np.float32)
        N.check(N.lib().fw_scene_linear_program(self._h, N.ptr(out), n))
        return out

    def walk_info(self):
        """{usable, top_meshes, top_wide_depth, mesh_wide_depth, prim_bits, triangles} (fw_scene_walk_info)."""
        out = np.zeros(6, np.int32)
        N.check(N.lib().fw_scene_walk_info(self._h, N.ptr(out)))
        return dict(zip(("usable", "top_meshes", "top_wide_depth", "mesh_wide_depth", "prim_bits", "triangles"), (int(v) for v in out)))

    def material_texture(self, material):
        return N.lib().fw_material_texture(self._h, material)

    def set_profiling(self, on: bool):
        N.check(N.lib().fw_set_profiling(self._h, 1 if on else 0))

    def set_batch_paths(self, paths: int):
        N.check(N.lib().fw_set_batch_paths(self._h, int(paths)))

    # ---- the hot path ----------------------------------------------------------------------------------
    def render(self, params: N.FwParams, want_rgb=True, want_sum=True):
        """fw_render with host buffers. Returns (rgb (H,W,3) u8 | None, sum (H,W,3) f32 | None, stats dict)."""
        h, w = params.height, params.width
        rgb = np.empty((h, w, 3), np.uint8) if want_rgb else None
        s = np.empty((h, w, 3), np.float32) if want_sum else None
        st = N.FwStats()
        N.check(N.lib().fw_render(self._h, C.byref(params), N.ptr(rgb) if want_rgb else None,
                                  N.ptr(s) if want_sum else None, C.byref(st)))
        return rgb, s, st.as_dict()

    def render_multi(self, params: N.FwParams, n_gpus: int, devices=None, reduce="nccl", want_rgb=True, want_sum=True):
        """fw_render_multi: the call's sample range split over `n_gpus` devices of this box (single process), sums
        combined on this scene's device by one NCCL reduce (`reduce="nccl"`) or by the fused peer-memory
        reduce + resolve kernel (`reduce="peer"`).  Returns (rgb | None, total sum | None, stats dict)."""
        h, w = params.height, params.width
        rgb = np.empty((h, w, 3), np.uint8) if want_rgb else None
        s = np.empty((h, w, 3), np.float32) if want_sum else None
        st = N.FwStats()
        ms = C.c_double()
        dev = None if devices is None else np.ascontiguousarray(devices, np.int32)
        N.check(N.lib().fw_render_multi(self._h, C.byref(params), int(n_gpus), None if dev is None else N.ptr(dev),
                                        {"nccl": 0, "peer": 1}[reduce], N.ptr(rgb) if want_rgb else None,
                                        N.ptr(s) if want_sum else None, C.byref(st), C.byref(ms)))
        d = st.as_dict()
        d["ms_reduce"] = ms.value
        return rgb, s, d

    def render_accumulate_device(self, params: N.FwParams, d_sum_ptr: int, stream_ptr: int = 0):
        """Adds the params' sample range into a device fp32 buffer (e.g. a torch tensor's data_ptr())."""
        st = N.FwStats()
        N.check(N.lib().fw_render_accumulate_device(self._h, C.byref(params), C.c_void_p(d_sum_ptr),
                                                    C.c_void_p(stream_ptr) if stream_ptr else None, C.byref(st)))
        return st.as_dict()

    def resolve_device(self, d_sum_ptr: int, npix: int, samples: int, gamma: float, d_rgb_ptr: int, stream_ptr: int = 0):
        N.check(N.lib().fw_resolve_device(self._h, C.c_void_p(d_sum_ptr), npix, samples, C.c_float(gamma),
                                          C.c_void_p(d_rgb_ptr), C.c_void_p(stream_ptr) if stream_ptr else None))

    # ---- probes ----------------------------------------------------------------------------------------
    def primary_rays(self, params: N.FwParams, sample: int, pix_begin=0, n=None):
        if n is None:
            n = params.width * params.height - pix_begin
        o = np.zeros((n, 3), np.float32)
        d = np.zeros((n, 3), np.float32)
        N.check(N.lib().fw_primary_rays(self._h, C.byref(params), sample, pix_begin, n, N.ptr(o), N.ptr(d)))
        return o, d

    def first_hit(self, origins, dirs, use_bvh: bool, seed=0, pixel=None, sample=None, bounce=None):
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        n = len(o)
        keys = [None if k is None else np.ascontiguousarray(k, np.uint32) for k in (pixel, sample, bounce)]
        res = {"obj": np.zeros(n, np.int32), "prim": np.zeros(n, np.int32), "material": np.zeros(n, np.int32),
               "t": np.zeros(n, np.float32), "point": np.zeros((n, 3), np.float32),
               "normal": np.zeros((n, 3), np.float32), "uv": np.zeros((n, 2), np.float32)}
        cnt = np.zeros(2, np.uint64)
        N.check(N.lib().fw_first_hit(self._h, 1 if use_bvh else 0, seed, n, N.ptr(o), N.ptr(d),
                                     *[None if k is None else N.ptr(k) for k in keys],
                                     N.ptr(res["obj"]), N.ptr(res["prim"]), N.ptr(res["material"]), N.ptr(res["t"]),
                                     N.ptr(res["point"]), N.ptr(res["normal"]), N.ptr(res["uv"]), N.ptr(cnt)))
        res["node_tests"], res["prim_tests"] = int(cnt[0]), int(cnt[1])
        return res

    def first_hit_wavefront(self, origins, dirs, use_bvh: bool, seed=0, sample=0, bounce=0):
        """The first-hit query through the kernels a render launches (fw_first_hit_wavefront); ray i is keyed as pixel i."""
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        n = len(o)
        res = {"obj": np.zeros(n, np.int32), "prim": np.zeros(n, np.int32), "material": np.zeros(n, np.int32),
               "t": np.zeros(n, np.float32), "point": np.zeros((n, 3), np.float32),
               "normal": np.zeros((n, 3), np.float32), "uv": np.zeros((n, 2), np.float32)}
        N.check(N.lib().fw_first_hit_wavefront(self._h, 1 if use_bvh else 0, seed, n, N.ptr(o), N.ptr(d), int(sample), int(bounce),
                                               N.ptr(res["obj"]), N.ptr(res["prim"]), N.ptr(res["material"]), N.ptr(res["t"]),
                                               N.ptr(res["point"]), N.ptr(res["normal"]), N.ptr(res["uv"])))
        return res

    def scatter_step(self, material, ray_o, ray_d, hit_t, hit_point, hit_normal, hit_uv, uniforms):
        n = len(material)
        material = np.ascontiguousarray(material, np.int32)
        arrs = [np.ascontiguousarray(a, np.float32) for a in (ray_o, ray_d, hit_t, hit_point, hit_normal, hit_uv)]
        uniforms = np.ascontiguousarray(uniforms, np.float32)
        res = {"emit": np.zeros((n, 3), np.float32), "scattered": np.zeros(n, np.int32),
               "atten": np.zeros((n, 3), np.float32), "o": np.zeros((n, 3), np.float32),
               "d": np.zeros((n, 3), np.float32), "consumed": np.zeros(n, np.int32)}
        N.check(N.lib().fw_scatter_step(self._h, n, N.ptr(material), *[N.ptr(a) for a in arrs], N.ptr(uniforms),
                                        uniforms.shape[1], N.ptr(res["emit"]), N.ptr(res["scattered"]),
                                        N.ptr(res["atten"]), N.ptr(res["o"]), N.ptr(res["d"]), N.ptr(res["consumed"])))
        return res

    def env_sample(self, dirs):
        d = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros_like(d)
        N.check(N.lib().fw_env_sample(self._h, len(d), N.ptr(d), N.ptr(out)))
        return out

    def texture_sample(self, tex, uv, point):
        uv = np.ascontiguousarray(uv, np.float32)
        pt = np.ascontiguousarray(point, np.float32)
        out = np.zeros_like(pt)
        N.check(N.lib().fw_texture_sample(self._h, int(tex), len(pt), N.ptr(uv), N.ptr(pt), N.ptr(out)))
        return out


def camera_constants(params: N.FwParams) -> np.ndarray:
    out = np.zeros(24, np.float32)
    N.check(N.lib().fw_camera(C.byref(params), N.ptr(out)))
    return out


def measure_peaks(device=0):
    f, l, sm, khz = C.c_double(), C.c_double(), C.c_int(), C.c_int()
    N.check(N.lib().fw_measure_peaks(device, C.byref(f), C.byref(l), C.byref(sm), C.byref(khz)))
    return {"fp32_tflops": f.value, "l2_gbs": l.value, "sm_count": sm.value, "sm_clock_khz": khz.value}


def selftest_shared_division(n_pairs: int, seed: int = 1, device: int = 0):
    """fw_selftest_shared_division: (bit mismatches, tiny-numerator threshold violations); both must be 0."""
    v = np.zeros(2, np.uint64)
    N.check(N.lib().fw_selftest_shared_division(device, C.c_uint64(n_pairs), C.c_uint64(seed), N.ptr(v)))
    return int(v[0]), int(v[1])


def texture_store_stats() -> dict:
    """fw_texture_store_stats: commits served from a resident array, uploads, arrays resident now and their bytes."""
    v = np.zeros(4, np.uint64)
    N.check(N.lib().fw_texture_store_stats(N.ptr(v)))
    return {"hits": int(v[0]), "uploads": int(v[1]), "arrays": int(v[2]), "bytes": int(v[3])}


def release_cached_memory():
    N.check(N.lib().fw_release_cached_memory())


def resolve_host(sums: np.ndarray, samples: int, gamma: float, device: int = 0) -> np.ndarray:
    """render.rs:184-189 for a host fp32 sum buffer (H, W, 3) -> u8 image, on the GPU (fw_resolve_host)."""
    s = np.ascontiguousarray(sums, np.float32)
    out = np.empty(s.shape, np.uint8)
    N.check(N.lib().fw_resolve_host(device, N.ptr(s), s.size // 3, int(samples), C.c_float(gamma), N.ptr(out)))
    return out


def render_scene(scene, renderer, device: int = 0):
    """`Renderer::render(scene)` on one GPU: YAML -> native scene -> fw_render. Returns (rgb, sum, stats)."""
    ns = NativeScene.from_scene(scene, device)
    try:
        return ns.render(renderer.params())
    finally:
        ns.close()
