/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the mlsgpu per-bin reconstruction core, used as the
 * parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (mlsgpu_b200/) never links, imports or calls anything in this directory.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * upstream mlsgpu tree).  Arithmetic conventions (they are the parity contract):
 *   - IEEE-754 binary32, round-to-nearest-even, no flush-to-zero;
 *   - built with -ffp-contract=off: a fused multiply-add happens only where fmaf() is
 *     written.  Where the OpenCL source leaves contraction to the compiler (mls.cl has
 *     FP_CONTRACT at its default ON) the choice made here is stated at the call site;
 *   - division and sqrt are correctly rounded; OpenCL's half_rsqrt(x), whose accuracy
 *     is implementation-defined, is taken as 1.0f / sqrtf(x).
 *
 * Parity pinning: checked against the reference's own known-answer tests
 * (test/test_splat_tree_cl.cpp, test/test_splat_tree.cpp, test/test_mls.cpp,
 * test/test_marching.cpp, test/test_mesh_filter.cpp) in tests/test_oracle_*.py.
 * The reference itself cannot be compiled in this environment (needs Boost, clogs,
 * an OpenCL runtime, cppunit: none present), so there is no oracle/_ref.
 */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/splat.h:40-45 and kernels/mls.cl:63-67: 32-byte splat. */
typedef struct
{
    float position[3];
    float radius;      /* radius on input; 1/radius^2 after the octree build */
    float normal[3];
    float quality;
} OrcSplat;

#define ORC_EXPORT __attribute__((visibility("default")))

static inline float orc_dot3(const float a[3], const float b[3])
{
    /* kernels/mls.cl:105-108 */
    return __builtin_fmaf(a[0], b[0], __builtin_fmaf(a[1], b[1], a[2] * b[2]));
}

static inline float orc_nan(void)
{
    union { uint32_t u; float f; } x;
    x.u = 0x7FC00000u;  /* nan(0U) */
    return x.f;
}

static inline int orc_isfinite(float x)
{
    union { uint32_t u; float f; } v;
    v.f = x;
    return (v.u & 0x7F800000u) != 0x7F800000u;
}

/* Stable LSD radix sort of (key,value) pairs on the low `bits` key bits (0 = all),
 * the published semantics of clogs::Radixsort (clogs 1.x; call sites
 * src/splat_tree_cl.cpp:308 and src/marching.cpp:572). */
uint32_t orc_make_code(int x, int y, int z);
void orc_decode(uint32_t code, int out[3]);

void orc_radixsort_u32_u32(uint32_t *keys, uint32_t *values, size_t n, unsigned bits);
void orc_radixsort_u64_f4(uint64_t *keys, float *values4, size_t n, unsigned bits);

#ifdef __cplusplus
}
#endif

#endif
