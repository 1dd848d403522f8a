// Seeded stand-in for the reference tests' init_seed() + random_number()
// (src/util.f90:72-102 seeds from the clock; a fixed seed keeps the C++
// restatements reproducible).
#pragma once
#include <cstdint>
struct rng64 {
    uint64_t s;
    explicit rng64(uint64_t seed) : s(seed) {}
    double next()   // uniform [0, 1)
    {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
};
