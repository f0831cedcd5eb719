// Shared helpers for libtrpx (sm_100a).  Error plumbing for the C ABI + small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <new>
#include "../../include/trpx.h"

namespace trpx {

// thread-local message returned by trpx_last_error()
char* err_buf();
int set_error(int code, const char* fmt, ...);

#define TRPX_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return ::trpx::set_error(TRPX_E_CUDA, "%s failed: %s (%s:%d)", #call,                    \
                                     cudaGetErrorString(e__), __FILE__, __LINE__);                   \
    } while (0)

#define TRPX_REQUIRE(cond, ...)                                                                      \
    do {                                                                                             \
        if (!(cond)) return ::trpx::set_error(TRPX_E_INVALID, __VA_ARGS__);                          \
    } while (0)

// RAII device switch: every entry point runs on its handle's device and restores the caller's.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            switched = true;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float gelu_new(float x) {
    // trphysx/transformer/utils.py:57-60 : 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))
    const float c = 0.7978845608028654f;
    float x3 = x * x * x;
    return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * x3)));
}

// The same function in sigmoid form: 0.5 x (1 + tanh(u)) == x / (1 + exp(-2u)).  One ex2 + one rcp instead of
// the ~25-instruction tanhf; relative deviation ~1e-7 (measured end-to-end in the parity tests).  Used by the
// rollout kernels, where the 128 gelu evaluations per trajectory-layer are a visible share of the issue slots.
__device__ __forceinline__ float gelu_new_sig(float x) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return __fdividef(x, 1.0f + __expf(-2.0f * u));
}

// Packed FP32 FMA (PTX fma.rn.f32x2, SASS FFMA2 with a scalar-broadcast multiplier): (d0, d1) += (a0, a1) * b.
// Same FP32 rate as two FFMAs (measured 74 vs 71 TFLOP/s, scripts/probes/ffma2_probe.cu) for ONE issue slot, and each
// element is an IEEE fused multiply-add, so results are bit-identical to the scalar form.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %4};\n mov.b64 rc, {%0, %1};\n"
        "fma.rn.f32x2 rc, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b));
}

// element-wise form: (d0, d1) += (a0, a1) * (b0, b1)
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%0, %1};\n"
        "fma.rn.f32x2 rc, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ float ex2_approx_fwd(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// gelu_new of two values at once with packed FP32 arithmetic (mul/fma/add.f32x2) + 2 ex2 + 2 rcp on the SFU:
// x * rcp(1 + 2^(k x (1 + 0.044715 x^2))), k = -2 log2(e) sqrt(2/pi) — the sigmoid form of gelu_new_sig, 4 issue
// slots per value instead of ~9.
__device__ __forceinline__ void gelu_new_sig2(float& x0, float& x1) {
    const float k = -2.0f * 1.4426950408889634f * 0.7978845608028654f;
    float e0, e1;
    asm("{\n.reg .b64 x, t, p, kk, ka;\n"
        "mov.b64 x, {%2, %3};\n"
        "mov.b64 kk, {%4, %4};\n"
        "mov.b64 ka, {%5, %5};\n"
        "mul.rn.f32x2 t, x, x;\n"                    // x^2
        "fma.rn.f32x2 p, t, ka, kk;\n"               // k (1 + 0.044715 x^2)
        "mul.rn.f32x2 p, p, x;\n"                    // exponent (log2 domain)
        "mov.b64 {%0, %1}, p;\n}"
        : "=f"(e0), "=f"(e1)
        : "f"(x0), "f"(x1), "f"(k), "f"(k * 0.044715f));
    e0 = ex2_approx_fwd(e0);
    e1 = ex2_approx_fwd(e1);
    float r0, r1;
    asm("{\n.reg .b64 e, one;\n"
        "mov.b64 e, {%2, %3};\n"
        "mov.b64 one, {%4, %4};\n"
        "add.rn.f32x2 e, e, one;\n"
        "mov.b64 {%0, %1}, e;\n}"
        : "=f"(r0), "=f"(r1)
        : "f"(e0), "f"(e1), "f"(1.0f));
    asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(r0));
    asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(r1));
    asm("{\n.reg .b64 x, r;\n"
        "mov.b64 x, {%0, %1};\n"
        "mov.b64 r, {%2, %3};\n"
        "mul.rn.f32x2 x, x, r;\n"
        "mov.b64 {%0, %1}, x;\n}"
        : "+f"(x0), "+f"(x1)
        : "f"(r0), "f"(r1));
}

// 2^x on the SFU (ex2.approx: max error 2 ulp over the whole range; inputs here are <= 0)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// cvt.rna.tf32.f32 (round to nearest, ties away, low 13 mantissa bits cleared) in two integer instructions: on sm_100a
// the PTX instruction compiles to FSETP + predicated IADD + LOP3 because it also passes Inf/NaN through unchanged.
// Every value split here is finite, for which this form is bit-identical (scripts/probes: SASS of cvt.rna.tf32).
__device__ __forceinline__ uint32_t tf32_rna_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// 256-bit global load (LDG.E.256 on sm_100a): one full 32-byte sector per lane.
__device__ __forceinline__ void ldg256(const float* p, float (&a)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
                 : "l"(p)
                 : "memory");
}

// Same with a compile-time byte offset folded into the instruction (no address register / LEA per row).
template <int OFF>
__device__ __forceinline__ void ldg256_off(const float* p, float (&a)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+%9];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
                 : "l"(p), "n"(OFF)
                 : "memory");
}

// Fire-and-forget bulk prefetch of a contiguous global range into L2 (cp.async.bulk.prefetch, SASS: UBLKPF).
// `bytes` must be a multiple of 16 and `p` 16-byte aligned.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- TMA bulk copy global -> shared with an mbarrier (cp.async.bulk, SASS: UBLKCP) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TRPX_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TRPX_DONE_%=;\n"
        "bra TRPX_WAIT_%=;\n"
        "TRPX_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Same wait for a thread that has nothing else to do for microseconds (a producer waiting for its consumers):
// back off between probes so that the spin does not eat the issue slots of the warps sharing its scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}

}  // namespace trpx
