// Stand-in for tiny_rng::Rng in the restated example scene generators (firework_b200/scenes.py SceneRng): the crate is not
// vendored with the reference, so both mirrors draw their scenes from the same SplitMix64 stream instead.
#pragma once
#include <cstdint>

struct SceneRng {
    uint64_t state;
    explicit SceneRng(uint64_t seed) : state(seed) {}
    uint64_t next_u64() {
        state += 0x9E3779B97F4A7C15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    float rand_f32() { return (float)((double)(next_u64() >> 40) * (1.0 / 16777216.0)); }
};
