// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers and small math utilities.
// Everything here is hand-written inline PTX for sm_100a (no CUTLASS dependency).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>

#define ARD_DEVINL __device__ __forceinline__

namespace ard {

// ------------------------------------------------------------------------------------------------ misc
ARD_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

ARD_DEVINL uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t.reg .b32 R1;\n\t"
        "elect.sync R1|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(pred));
    return pred;
}

ARD_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Programmatic dependent launch (enqueue_pdl in ard_internal.h): a kernel launched with the programmatic-serialization attribute
// may START while its predecessor in the stream is still running. pdl_launch_dependents() (first statement of every kernel)
// lets the successor's CTAs be scheduled as soon as SMs free up; pdl_wait() blocks until the predecessor grid has completed
// and its writes are visible, and must precede the first access to any global buffer (reads AND writes: the predecessor may
// still be reading what this kernel overwrites). Everything before it - barrier / TMEM / shared-memory set-up, descriptor
// prefetch - overlaps the predecessor's tail. Both are no-ops in a kernel launched without the attribute.
ARD_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
ARD_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ mbarrier
ARD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
ARD_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ARD_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

ARD_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ARD_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
ARD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
ARD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Wait for roles that expect to idle for thousands of cycles (epilogue / issuer warps of the fused kernels). A bare
// try_wait returns after ~20 cycles, so every idle warp keeps issuing a TRYWAIT + BRA pair: in the fused FFN half of all
// issued instructions were such polls, taken from the schedulers the GELU / LayerNorm warps were running on. The suspend-time
// hint lets the hardware park the warp until the phase completes (or the hint expires) instead.
ARD_DEVINL void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ TMA
ARD_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
ARD_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
ARD_DEVINL void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
ARD_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
ARD_DEVINL void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
ARD_DEVINL void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------ tcgen05 / TMEM
ARD_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
ARD_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
ARD_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
ARD_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ARD_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> f32, issued by ONE thread.
ARD_DEVINL void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A run of NK (k-step) MMAs over one swizzled k-block in ONE asm statement: D (+)= A[:, 16k..] * B[:, 16k..], descriptors
// advance by 2 (32 bytes) per k-step. Issued from divergent code (lane 0 of the issuer warp) every tcgen05.mma costs a dozen
// SASS instructions when written as separate statements (R2UR of each operand + an ELECT / BRA.U.ANY wrapper per
// instruction, ~85 cycles per MMA measured in the fused FFN); inside one statement the operands are converted once.
// `first_acc` = accumulate flag of the first MMA (the rest always accumulate).
#define ARD_UMMA_RUN2(CG)                                                                      \
    asm volatile(                                                                              \
        "{\n\t.reg .pred p;\n\t.reg .b64 a1, b1;\n\t"                                         \
        "setp.ne.b32 p, %4, 0;\n\t"                                                            \
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"                                         \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], %1, %2, %3, p;\n\t"                      \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], a1, b1, %3, 1;\n\t}\n"                   \
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(first_acc)                    \
        : "memory")
#define ARD_UMMA_RUN4(CG)                                                                      \
    asm volatile(                                                                              \
        "{\n\t.reg .pred p;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"                         \
        "setp.ne.b32 p, %4, 0;\n\t"                                                            \
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"                                         \
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"                                         \
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"                                         \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], %1, %2, %3, p;\n\t"                      \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], a1, b1, %3, 1;\n\t"                      \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], a2, b2, %3, 1;\n\t"                      \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], a3, b3, %3, 1;\n\t}\n"                   \
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(first_acc)                    \
        : "memory")
template <int NK, bool PAIR = false>
ARD_DEVINL void umma_f16_ss_run(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t first_acc) {
    static_assert(NK == 2 || NK == 4, "k-steps per run: 2 (SWIZZLE_64B k-block / half of a 128B one) or 4 (SWIZZLE_128B)");
    if constexpr (NK == 2) {
        if constexpr (PAIR) ARD_UMMA_RUN2("2"); else ARD_UMMA_RUN2("1");
    } else {
        if constexpr (PAIR) ARD_UMMA_RUN4("2"); else ARD_UMMA_RUN4("1");
    }
}
// Same run with the A operand in TENSOR MEMORY (M = 128: row i in lane i, two 16-bit K elements per 32-bit column, so a
// K = 16 step is 8 columns): D (+)= A_tmem[:, 16k..] * B[:, 16k..]. Four k-steps = one 64-wide k-block.
ARD_DEVINL void umma_f16_ts_run4(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t first_acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 b1, b2, b3;\n\t.reg .b32 a1, a2, a3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.s64 b1, %2, 2;\n\tadd.s64 b2, %2, 4;\n\tadd.s64 b3, %2, 6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, 1;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(first_acc)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
ARD_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
ARD_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Arrives (count 1) on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
// Implies tcgen05.fence::before_thread_sync.
ARD_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster (same TPC) run one M=256 UMMA; the even-ranked CTA leads.
// A shared::cta address with bit 24 cleared names the same offset in the LEADER CTA's shared memory (shared::cluster window).
constexpr uint32_t PAIR_LEADER_MASK = 0xFEFFFFFFu;
ARD_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
ARD_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into THIS CTA's shared memory whose transaction bytes are counted on the LEADER CTA's mbarrier.
ARD_DEVINL void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PAIR_LEADER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
ARD_DEVINL void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
ARD_DEVINL void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
ARD_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D (256 x N: rows 0-127 in the leader's TMEM, 128-255 in the peer's) (+)= A (each CTA's 128 x 16 tile) * B (N x 16, each CTA
// holds N/2 rows). Issued by ONE thread of the leader CTA.
ARD_DEVINL void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier at this offset in BOTH CTAs of the pair once the previously issued pair MMAs have completed.
ARD_DEVINL void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// Arrive on the LEADER CTA's copy of `bar` (from either CTA of the pair), cluster-scope release.
ARD_DEVINL void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PAIR_LEADER_MASK) : "memory");
}
ARD_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread i gets lane (warp%4)*32+i).
ARD_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
ARD_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
ARD_DEVINL void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
ARD_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes (64 bf16),
// 8-row groups 1024 B apart (SBO), tile base 1024-byte aligned. (cute::UMMA::SmemDescriptor: start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).)
ARD_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// Same for CU_TENSOR_MAP_SWIZZLE_64B tiles: rows of 64 bytes (32 bf16), 8-row groups 512 B apart, base 512-byte aligned.
ARD_DEVINL uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;        // SBO
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b_format BF16 [7,10)/[10,13)=1,
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same with fp16 A and B operands (a_format = b_format = F16 = 0), fp32 accumulation.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ math
// erf(x) for the exact-erf GELU of Mlp (htsat.py:151 act_layer=nn.GELU):
//   erf(x) = sign(x) * (1 - 2^(-|x| Q(|x|))),  Q = degree-5 fit of -log2(erfc(t))/t on (0, 3.9]  (tools/fit_erf.py)
// max |error| 3.4e-7 in float32 (GELU abs error 2.8e-7); one MUFU.EX2, no branches.
ARD_DEVINL float erf_fast(float x) {
    float t = fminf(fabsf(x), 3.9f);
    float q = -1.593649553e-04f;
    q = fmaf(q, t, 3.748819894e-03f);
    q = fmaf(q, t, -3.104246184e-02f);
    q = fmaf(q, t, 1.498086252e-01f);
    q = fmaf(q, t, 9.181317066e-01f);
    q = fmaf(q, t, 1.627928301e+00f);
    float r = 1.0f - exp2f(-q * t);
    return copysignf(r, x);
}
// d/dx gelu(x) = Phi(x) + x phi(x) = 0.5 (1 + erf(x / sqrt 2)) + x exp(-x^2 / 2) / sqrt(2 pi)
ARD_DEVINL float gelu_erf_grad(float x) {
    const float cdf = fmaf(0.5f, erf_fast(x * 0.70710678118654752f), 0.5f);
    const float pdf = 0.3989422804014327f * exp2f(-0.72134752044448170f * x * x);
    return fmaf(x, pdf, cdf);
}
// erf GELU (Mlp act_layer=nn.GELU, htsat.py:151) of two values in packed fp16 arithmetic: the fp32 pre-activations are
// rounded to half2 once and the result stays fp16 - it is the A operand of the fp16 fc2 GEMM. (10-bit mantissa) 5x closer
// to the exact GELU than the fp32-GELU -> bf16 operand path (rel. l2 error 3.2e-4 vs 1.7e-3 on N(0,1.5) inputs,
// emulated in tools/fit_erf.py --f16).
//
// Formula: erf(x / sqrt 2) = tanh(x (c0 + c1 x^2)), the classic tanh form with c0, c1 re-fitted to the ERF GELU (minimax on
// the absolute GELU error, tools/fit_erf.py --tanh): max |GELU error| 2.7e-4, below fp16 resolution of the values that matter;
// evaluated in fp16 the rel. l2 error against the float64 erf GELU is 3.7e-4 on N(0,1.5) inputs (3.2e-4 for the
// 1 - 2^(-tQ(t)) form it replaces). The argument is y = x / 2 (the caller folds the halving into its fp32 bias add:
// y = 0.5 acc + 0.5 b): gelu(x) = y + y tanh(y (2 c0 + 8 c1 y^2)), 4 packed ops + 2 MUFU.TANH.F16 + PRMT per pair. Measured
// on B200 (tools/micro/mufu_rate.cu): HFMA2 issues at 64 lanes/clk/SM - half the FFMA rate - and MUFU at 16 elements/clk/SM
// whatever the format, so the cost of a packed GELU is its op count: the previous form's h2exp2 expanded to 2 cvt + 2 MUFU.EX2
// + 2 FFMA + pack and its polynomial was 6 HFMA2 (25 instructions per pair; now 9). No clamp is needed: y^2 -> inf gives
// p = +inf, u = +-inf, tanh = +-1.
ARD_DEVINL uint32_t gelu_erf_f16x2_halved(float ya, float yb) {
    const __half2 y = __floats2half2_rn(ya, yb);
    const __half2 p = __hfma2(__float2half2_rn(8.0f * 3.470089e-02f), __hmul2(y, y), __float2half2_rn(2.0f * 8.0015708e-01f));
    const __half2 u = __hmul2(y, p);
    uint32_t tb;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
    const __half2 g = __hfma2(y, *reinterpret_cast<const __half2*>(&tb), y);
    return *reinterpret_cast<const uint32_t*>(&g);
}
// d/dx of the erf GELU for two pre-activations held as a packed bf16 pair (the FFN backward: dh = (g W2) * gelu'(hpre)), in packed
// fp16 arithmetic:  gelu'(x) = 1/2 + t/2 + (1 - t^2) x (d0 + d1 x^2) / 2,  t = tanh(x (c0 + c1 x^2)) - the derivative shape of the
// tanh form with all four constants fitted to the exact Phi(x) + x phi(x) (minimax, tools/fit_erf.py --grad: max |error| 1.2e-4;
// evaluated in fp16 on bf16-rounded N(0, 1.5) inputs rel. l2 3.7e-4, less than the 5e-4 the bf16 rounding of hpre itself costs).
// 10 packed ops + 1 tanh.approx.f16x2 per pair against ~25 fp32 instructions + 2 MUFU per ELEMENT for gelu_erf_grad: the gelu'
// GEMM epilogue was issue-bound at 2.4 elements/clk/SM. x^2 is clamped at 64 (t = +-1 there) so large |x| gives 0 * finite.
ARD_DEVINL float2 gelu_erf_grad_h2(__half2 x) {
    const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(64.0f));
    const __half2 u = __hmul2(x, __hfma2(__float2half2_rn(2.934764e-02f), x2, __float2half2_rn(6.9744092e-01f)));
    uint32_t tb;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
    const __half2 t = *reinterpret_cast<const __half2*>(&tb);
    const __half2 w = __hmul2(x, __hfma2(__float2half2_rn(-0.5f * 1.019215e-02f), x2, __float2half2_rn(0.5f * 8.9857494e-01f)));
    const __half2 s = __hfma2(__hneg2(t), t, __float2half2_rn(1.0f));
    const __half2 r = __hfma2(t, __float2half2_rn(0.5f), __float2half2_rn(0.5f));
    return __half22float2(__hfma2(s, w, r));
}
ARD_DEVINL float2 gelu_erf_grad_bf16x2(uint32_t hb) {   // the two pre-activations as a packed bf16 pair
    return gelu_erf_grad_h2(__floats2half2_rn(__uint_as_float(hb << 16), __uint_as_float(hb & 0xffff0000u)));
}
ARD_DEVINL float gelu_erf(float x) {
    float h = 0.5f * x;
    return fmaf(h, erf_fast(x * 0.70710678118654752f), h);
}

}  // namespace ard
