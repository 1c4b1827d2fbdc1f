// PTX helpers shared by the convolution kernels (sm_100a): TF32 conversion, mma.sync, cp.async, mbarrier, TMA bulk copy,
// tcgen05 (TMEM allocation, UMMA issue / commit, TMEM loads) and the rule lookup.
#pragma once
#include "geometry.cuh"

namespace mopa {

// ------------------------------------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz));
}
// arrive on `bar` (without incrementing its pending count) once all cp.async copies this thread has issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ int gather_lookup(const Gather &g, int k, int64_t row) {
    if (g.table) return __ldg(g.table + (int64_t)k * g.ld + row);
    return (__ldg(g.kidx + row) == k) ? __ldg(g.parent + row) : -1;
}

// ------------------------------------------------------------------------------------------------ mbarrier / TMA helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint (ns): the hardware parks the warp until the phase completes or the hint expires, so a
// waiting warp re-issues the instruction every ~10 ms instead of spinning (ncu: the unhinted loops were ~14 % of all
// executed instructions of k_conv_tc, competing for issue slots with the warps that do the work)
constexpr uint32_t kMbarSuspendNs = 10000000u;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


// ------------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t *bar) { mbar_arrive(bar); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// whole warp: allocate `cols` (power of two >= 32) TMEM columns, base address written to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// true on exactly one lane of a converged warp; lets ptxas keep the surrounding (warp-uniform) values in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, fp32 accumulate; issued by ONE thread (SASS: UTCHMMA-family)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with a per-output-row write mask: bit r of {m0..m3} set = row r of D (TMEM lane r) is NOT updated (PTX
// "disable-output-lane"); always accumulates. The masks travel in four consecutive uniform registers (4 R2UR per call site).
__device__ __forceinline__ void umma_tf32_masked(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
// zero 16 consecutive fp32 columns of this thread's TMEM lane (SASS: STTM); complete after tmem_st_wait()
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor: K-major operand, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: TF32 x TF32 -> F32, both operands K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// shared-memory accesses by 32-bit shared-window address (keeps ptxas from falling back to generic LD/ST + 64-bit math)
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async16_s(uint32_t dst, const void *gmem_src, int src_bytes) {  // src_bytes 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
}
// predicated form: compiles to one @P LDGSTS (an `if` around the asm costs a branch + reconvergence pair per site)
__device__ __forceinline__ void cp_async16_pred_s(uint32_t dst, const void *gmem_src, int src_bytes, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}" ::"r"(dst),
        "l"(gmem_src), "r"(src_bytes), "r"((int)pred));
}
// guarded copy with the ignore-src operand: @guard { dst[0..16) = zero ? 0 : src[0..16) }. One LDGSTS(.ZFILL) and two
// predicate moves; the src-size register form makes ptxas emit address arithmetic around every copy.
__device__ __forceinline__ void cp_async16_zfill_pred_s(uint32_t dst, const void *gmem_src, uint32_t guard, uint32_t zero) {
    asm volatile(
        "{\n\t.reg .pred pg, pz;\n\tsetp.ne.b32 pg, %2, 0;\n\tsetp.ne.b32 pz, %3, 0;\n\t"
        "@pg cp.async.cg.shared.global [%0], [%1], 16, pz;\n\t}" ::"r"(dst),
        "l"(gmem_src), "r"(guard), "r"(zero));
}
// guarded plain copy: @guard dst[0..16) = src[0..16). One @P LDGSTS.
__device__ __forceinline__ void cp_async16_guard_s(uint32_t dst, const void *gmem_src, uint32_t guard) {
    asm volatile("{\n\t.reg .pred pg;\n\tsetp.ne.b32 pg, %2, 0;\n\t@pg cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst),
                 "l"(gmem_src), "r"(guard));
}
// unguarded copy with the ignore-src operand: dst[0..16) = zero ? 0 : src[0..16)
__device__ __forceinline__ void cp_async16_zfill_s(uint32_t dst, const void *gmem_src, uint32_t zero) {
    asm volatile("{\n\t.reg .pred pz;\n\tsetp.ne.b32 pz, %2, 0;\n\tcp.async.cg.shared.global [%0], [%1], 16, pz;\n\t}" ::"r"(dst),
                 "l"(gmem_src), "r"(zero));
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc_s(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void named_barrier_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_upto() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace mopa
