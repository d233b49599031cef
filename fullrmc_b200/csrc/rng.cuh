// rng.cuh -- the counter-based random numbers of device-generated runs of moves.
//
// The contract is defined in fullrmc_b200/rng.py (Philox4x32-10, key = the 64-bit seed, counter = (step lo, step hi,
// block, 0); what each word of a step is used for; the float32 operation order of the translation vector, which
// follows generate_random_vector, Core/Collection.py:674-701, and of transform_coordinates,
// Extensions/boundary_conditions_collection.pyx:88-110).  This header repeats it operation by operation so that the
// device draws bit for bit what a reference Engine with the plug-ins of fullrmc_b200/engine_plugins.py draws.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace frmc {

struct PhiloxWords { uint32_t w[4]; };

__host__ __device__ __forceinline__ PhiloxWords philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    PhiloxWords out;
    out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
    return out;
}

// (w >> 8) * 2^-24: a float32 in [0, 1), exact
__device__ __forceinline__ float philox_uniform(uint32_t w) { return __fmul_rn((float)(w >> 8), 5.9604644775390625e-08f); }

struct StepRandom {
    uint32_t group_word;      // block 0 word 0: group = (w * numberOfGroups) >> 32
    float vx, vy, vz;         // translation vector
    float accept;             // this step's generate_random_float()
};

__device__ __forceinline__ StepRandom step_random(uint64_t seed, uint64_t counter, float min_amp, float max_amp)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), lo = (uint32_t)counter, hi = (uint32_t)(counter >> 32);
    const PhiloxWords a = philox4x32_10(lo, hi, 0u, 0u, k0, k1), b = philox4x32_10(lo, hi, 1u, 0u, k0, k1);
    StepRandom s;
    s.group_word = a.w[0];
    float vx = __fsub_rn(1.0f, __fmul_rn(2.0f, philox_uniform(a.w[1])));
    float vy = __fsub_rn(1.0f, __fmul_rn(2.0f, philox_uniform(a.w[2])));
    float vz = __fsub_rn(1.0f, __fmul_rn(2.0f, philox_uniform(a.w[3])));
    float n2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
    if (n2 == 0.0f) { vx = 1.0f; vy = 0.0f; vz = 0.0f; n2 = 1.0f; }
    const float norm = __fsqrt_rn(n2);
    vx = __fdiv_rn(vx, norm); vy = __fdiv_rn(vy, norm); vz = __fdiv_rn(vz, norm);
    const float amp = __fmul_rn(philox_uniform(b.w[0]), __fsub_rn(max_amp, min_amp));
    s.vx = __fadd_rn(__fmul_rn(vx, amp), __fmul_rn(vx, min_amp));
    s.vy = __fadd_rn(__fmul_rn(vy, amp), __fmul_rn(vy, min_amp));
    s.vz = __fadd_rn(__fmul_rn(vz, amp), __fmul_rn(vz, min_amp));
    s.accept = philox_uniform(b.w[1]);
    return s;
}

// transform_coordinates (boundary_conditions_collection.pyx:104-107): out_k = (x*m[0][k] + y*m[1][k]) + z*m[2][k]
__device__ __forceinline__ void transform_point(const float *__restrict__ m, float x, float y, float z, float &ox, float &oy, float &oz)
{
    ox = __fadd_rn(__fadd_rn(__fmul_rn(x, m[0]), __fmul_rn(y, m[3])), __fmul_rn(z, m[6]));
    oy = __fadd_rn(__fadd_rn(__fmul_rn(x, m[1]), __fmul_rn(y, m[4])), __fmul_rn(z, m[7]));
    oz = __fadd_rn(__fadd_rn(__fmul_rn(x, m[2]), __fmul_rn(y, m[5])), __fmul_rn(z, m[8]));
}

}  // namespace frmc
