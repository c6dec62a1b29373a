// K4-SELL — the ring kernel's stage format for value-indexed plans: sliced ELLPACK inside every block.
// (included by spmv.cu after the ring kernel; uses its TMA / mbarrier helpers, DotArgs and the halo protocol)
//
// Why: with 8-bit value codes and 16-bit window positions a stage carries 3 bytes per non-zero and the kernel stops being
// HBM-bound — the consumers' shared-memory loads become the limit (ncu, profiles/ncu_r02_v8.raw.csv: LSU data pipe 86 % busy,
// a third of its wavefronts bank conflicts: one thread per row walks its entries with a stride of row-length elements, and
// strides of 7 bytes / 7 half-words collide where 7 words did not).  So the plan lays the compressed entries of a block out
// the way the consumers read them: rows are taken 32 at a time (a slice = one warp), entry j of the slice's rows is stored
// at slice_base + 32 j + lane.  A warp then reads 32 consecutive codes (one wavefront), 32 consecutive positions (one
// wavefront), gathers x from the staged windows and looks the values up in the block's dictionary; rows shorter than the
// slice's longest are padded (the padded entries are skipped, not multiplied by zero: -0 and NaN stay exact).  No row
// offsets are needed, only one length byte per row and one offset word per slice.
//
// Per row the operands, their order and the roundings are the reference's (sparsematrix.rs:146-158): bit-identical results.
// The plan keeps it only if the padding stays below 20 % (stencils, regular meshes); ragged matrices keep the CRS-order stage.
#pragma once

namespace smb {

struct alignas(16) SellBlock {            // one per row block; the producer loads it, the consumers get the first 32 bytes
    unsigned long long r0;                // first row of the block
    unsigned rows, n_slices;              // rows, slices of 32 rows
    unsigned e_bytes;                     // padded entries of the block (= bytes of codes; positions take twice as much)
    unsigned flags;                       // kSellGhost: a window reaches past g0 (distributed plans); kSellSelf + position << 8: see below
    unsigned long long ebase;             // first padded entry of the block in s_codes / s_cols (a multiple of 32)
    unsigned long long rbase;             // first row-length byte (a multiple of 16)
    unsigned long long sbase;             // first slice-offset word (a multiple of 4)
    unsigned long long lo[kNSeg];         // x windows
    unsigned len[kNSeg];
};
static_assert(sizeof(SellBlock) == 96, "SellBlock layout");
constexpr unsigned kSellGhost = 1u;
// kSellSelf: one staged x window covers the block's own rows [r0, r0 + rows) as columns; flags >> 8 is the element position
// of x[r0] inside the staged windows.  The fused dot x.(A x) of a CG iteration then takes its weight from shared memory
// instead of reading x a second time from HBM.
constexpr unsigned kSellSelf = 2u;

struct SellDesc {                          // what the consumers need of a staged block
    unsigned long long r0;
    unsigned rows, n_slices;
    unsigned self_pos;                     // position of x[r0] in the staged windows, ~0u: not staged
};

// ---- plan time -----------------------------------------------------------------------------------------------------
// Pass 1, one warp per slice of 32 rows: the slice's width (its longest row).  widths[first_slice_of_block + q].
template <class I>
__global__ void sell_width_kernel(const I* __restrict__ offs, const I* __restrict__ blk_rows, const unsigned long long* __restrict__ slice_base,
                                  unsigned long long n_blocks, unsigned* __restrict__ widths) {
    const unsigned long long b = blockIdx.x;
    const unsigned long long r0 = (unsigned long long)blk_rows[b], r1 = (unsigned long long)blk_rows[b + 1];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const unsigned n_slices = (unsigned)((r1 - r0 + 31) / 32);
    for (unsigned q = warp; q < n_slices; q += n_warps) {
        const unsigned long long r = r0 + 32ull * q + lane;
        unsigned len = r < r1 ? (unsigned)((unsigned long long)offs[r + 1] - (unsigned long long)offs[r]) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        if (lane == 0) widths[slice_base[b] + q] = len;
    }
}

// Per block: padded entries, padded rows, padded slice words -> three arrays that are then scanned into bases.
__global__ void sell_block_sizes_kernel(const unsigned* __restrict__ widths, const unsigned long long* __restrict__ slice_base,
                                        const unsigned long long* __restrict__ rows_of, unsigned long long n_blocks,
                                        unsigned long long* __restrict__ e_of, unsigned long long* __restrict__ rpad_of,
                                        unsigned long long* __restrict__ spad_of) {
    const unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const unsigned long long rows = rows_of[b], ns = (rows + 31) / 32;
    unsigned long long e = 0;
    for (unsigned long long q = 0; q < ns; ++q) e += 32ull * widths[slice_base[b] + q];
    e_of[b] = e;
    rpad_of[b] = (rows + 15) & ~15ull;
    spad_of[b] = (ns + 1 + 3) & ~3ull;                 // n_slices + 1 offsets
}

// Pass 2, one CTA per block: slice offsets, row lengths, and the entries permuted from CRS order (vcodes / lcols, both indexed
// k - codes_base) into the slices.  Padding stays zero (the arrays are cleared before).
template <class I>
__global__ void __launch_bounds__(256)
sell_fill_kernel(const I* __restrict__ offs, const I* __restrict__ blk_rows, const unsigned* __restrict__ widths,
                 const unsigned long long* __restrict__ slice_base, const unsigned long long* __restrict__ ebase,
                 const unsigned long long* __restrict__ rbase, const unsigned long long* __restrict__ sbase,
                 const uint8_t* __restrict__ vcodes, const uint16_t* __restrict__ lcols, unsigned long long codes_base, unsigned vbytes,
                 const unsigned long long* __restrict__ seg_lo, const unsigned* __restrict__ seg_len, unsigned long long g0,
                 uint8_t* __restrict__ s_codes, uint16_t* __restrict__ s_cols, uint8_t* __restrict__ s_rowlen, unsigned* __restrict__ s_soff,
                 SellBlock* __restrict__ blocks) {
    __shared__ unsigned soff[2049];                   // (a block holds at most 65536 rows: 2048 slices)
    const unsigned long long b = blockIdx.x;
    const unsigned long long r0 = (unsigned long long)blk_rows[b], r1 = (unsigned long long)blk_rows[b + 1];
    const unsigned rows = (unsigned)(r1 - r0), n_slices = (rows + 31) / 32;
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (unsigned q = 0; q < n_slices; ++q) { soff[q] = run; run += 32u * widths[slice_base[b] + q]; }
        soff[n_slices] = run;
        SellBlock sb;
        memset(&sb, 0, sizeof sb);
        sb.r0 = r0; sb.rows = rows; sb.n_slices = n_slices; sb.e_bytes = run;
        sb.ebase = ebase[b]; sb.rbase = rbase[b]; sb.sbase = sbase[b];
        bool ghost = false;
        unsigned at = 0, self = ~0u;
        for (int i = 0; i < kNSeg; ++i) {
            sb.lo[i] = seg_lo[kNSeg * b + i]; sb.len[i] = seg_len[kNSeg * b + i];
            ghost = ghost || (sb.len[i] != 0u && sb.lo[i] + sb.len[i] > g0);
            if (self == ~0u && sb.len[i] != 0u && sb.lo[i] <= r0 && r1 <= sb.lo[i] + sb.len[i] && r1 <= g0) self = at + (unsigned)(r0 - sb.lo[i]);
            at += sb.len[i];
        }
        sb.flags = (ghost ? kSellGhost : 0u) | (self != ~0u && self < (1u << 24) ? kSellSelf | (self << 8) : 0u);
        blocks[b] = sb;
    }
    __syncthreads();
    for (unsigned q = threadIdx.x; q <= n_slices; q += 256) s_soff[sbase[b] + q] = soff[q];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (unsigned q = warp; q < n_slices; q += 8) {
        const unsigned long long r = r0 + 32ull * q + lane;
        if (r >= r1) continue;
        const unsigned long long ka = (unsigned long long)offs[r], ke = (unsigned long long)offs[r + 1];
        s_rowlen[rbase[b] + 32ull * q + lane] = (uint8_t)(ke - ka);
        const unsigned long long dst = ebase[b] + soff[q] + lane;
        for (unsigned long long k = ka; k < ke; ++k) {
            s_codes[dst + 32ull * (k - ka)] = vcodes[k - codes_base];
            // the window position as a BYTE offset into the staged x windows (they live inside a 55 KB stage, so it fits 16 bits):
            // the consumers' gather then needs no address arithmetic
            s_cols[dst + 32ull * (k - ka)] = (uint16_t)((unsigned)lcols[k - codes_base] * vbytes);
        }
    }
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// Same ring as spmv_ring_kernel: a producer thread fills `stages` stages by bulk copies (codes, positions, row lengths,
// slice offsets, x windows, dictionary), 15 consumer warps drain them slice by slice.
template <class T, bool DOT, bool DIST>
__global__ void __launch_bounds__(kRingThreads, 2)
spmv_ring_sell_kernel(const SellBlock* __restrict__ blocks, const uint8_t* __restrict__ s_codes, const uint16_t* __restrict__ s_cols,
                      const uint8_t* __restrict__ s_rowlen, const unsigned* __restrict__ s_soff, const T* __restrict__ vdict,
                      unsigned n_blocks, unsigned ecap, unsigned rcap, unsigned scap, unsigned xcap, unsigned stages,
                      const T* __restrict__ x, T* __restrict__ y, DotArgs dot, const typename HaloParam<DIST>::type halo, unsigned rot) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kPipeMaxStages];
    __shared__ __align__(8) uint64_t empty[kPipeMaxStages];
    __shared__ SellDesc s_desc[kPipeMaxStages];
    constexpr unsigned kConsumerWarps = kRingThreads / 32 - 1;
    const unsigned tid = threadIdx.x;
    // stage: [codes ecap][positions 2 ecap][row lengths rcap][slice offsets 4 scap][x windows xcap T][dictionary 256 T]
    const size_t o_cols = (size_t)ecap;
    const size_t o_len = o_cols + 2 * (size_t)ecap;
    const size_t o_soff = o_len + (size_t)rcap;
    const size_t o_x = o_soff + 4 * (size_t)scap;
    const size_t o_dict = o_x + (size_t)xcap * sizeof(T);
    const size_t stage_bytes = o_dict + 256 * sizeof(T);
    double stop = 0.0;
    if constexpr (DOT) { if (dot.done != nullptr) stop = __ldcg(dot.done); }
    if (stop != 0.0) return;
    const unsigned n_my = blockIdx.x < n_blocks ? (n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0)
        for (unsigned s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    unsigned long long epoch = 0, g0 = ~0ull;
    const T* gx = nullptr;
    if constexpr (DIST) {
        epoch = __ldcg(halo.epoch) + 1ull;
        g0 = halo.g0;
        gx = (const T*)halo.ghost + (epoch & 1ull) * halo.ghost_stride;
    }
    [[maybe_unused]] bool waited = false;
    __syncthreads();
    double acc = 0.0;

    if (tid < 32) {
        if (tid == 0) {
            unsigned s = 0, parity = 0;
            for (unsigned j = 0; j < n_my; ++j) {
                size_t b = (size_t)blockIdx.x + (size_t)j * gridDim.x;
                if constexpr (DIST) { b += rot; if (b >= n_blocks) b -= n_blocks; }
                const SellBlock* sb = blocks + b;
                const ulonglong2 h0 = __ldg(reinterpret_cast<const ulonglong2*>(sb));                 // r0 | rows, n_slices
                const uint4 h1 = __ldg(reinterpret_cast<const uint4*>(sb) + 1);                       // e_bytes, flags, ebase
                const ulonglong2 h2 = __ldg(reinterpret_cast<const ulonglong2*>(sb) + 2);             // rbase, sbase
                const ulonglong2 l0 = __ldg(reinterpret_cast<const ulonglong2*>(sb) + 3);             // lo[0..1]
                const ulonglong2 l1 = __ldg(reinterpret_cast<const ulonglong2*>(sb) + 4);             // lo[2..3]
                const uint4 ln = __ldg(reinterpret_cast<const uint4*>(sb) + 5);                       // len[0..3]
                const unsigned rows = (unsigned)(h0.y & 0xffffffffull), n_slices = (unsigned)(h0.y >> 32);
                const unsigned e_bytes = h1.x, flags = h1.y;
                const unsigned long long ebase = (unsigned long long)h1.z | ((unsigned long long)h1.w << 32);
                const unsigned long long lo[kNSeg] = {l0.x, l0.y, l1.x, l1.y};
                const unsigned len[kNSeg] = {ln.x, ln.y, ln.z, ln.w};
                const unsigned rbytes = (rows + 15u) & ~15u, sbytes = 4u * ((n_slices + 1u + 3u) & ~3u);
                unsigned xtotal = 0;
#pragma unroll
                for (int i = 0; i < kNSeg; ++i) xtotal += len[i];
                if (j >= stages) mbar_wait(&empty[s], parity ^ 1u);
                unsigned char* base = smem_raw + (size_t)s * stage_bytes;
                s_desc[s].r0 = h0.x; s_desc[s].rows = rows; s_desc[s].n_slices = n_slices;
                if constexpr (DOT) s_desc[s].self_pos = (flags & kSellSelf) != 0u && dot.w == (const void*)x ? flags >> 8 : ~0u;
                if constexpr (DIST) {
                    if (!waited && ((flags & kSellGhost) != 0u || (blockIdx.x == 0 && j + 1 == n_my))) {
                        halo_wait(halo, epoch);
                        fence_proxy_async_global();
                        waited = true;
                    }
                }
                mbar_expect_tx(&full[s], 3u * e_bytes + rbytes + sbytes + xtotal * (unsigned)sizeof(T) + 256u * (unsigned)sizeof(T));
                if (e_bytes) {
                    bulk_g2s(base, s_codes + ebase, e_bytes, &full[s]);
                    bulk_g2s(base + o_cols, s_cols + ebase, 2u * e_bytes, &full[s]);
                }
                bulk_g2s(base + o_len, s_rowlen + h2.x, rbytes, &full[s]);
                bulk_g2s(base + o_soff, s_soff + h2.y, sbytes, &full[s]);
                bulk_g2s(base + o_dict, vdict + 256 * b, 256u * (unsigned)sizeof(T), &full[s]);
                unsigned at = 0;
#pragma unroll
                for (int i = 0; i < kNSeg; ++i)
                    if (len[i]) {
                        unsigned char* dstw = base + o_x + (size_t)at * sizeof(T);
                        if constexpr (!DIST) {
                            bulk_g2s(dstw, x + lo[i], len[i] * (unsigned)sizeof(T), &full[s]);
                        } else {
                            const unsigned n_own = lo[i] >= g0 ? 0u : (unsigned)(g0 - lo[i] < (unsigned long long)len[i] ? g0 - lo[i] : len[i]);
                            if (n_own) bulk_g2s(dstw, x + lo[i], n_own * (unsigned)sizeof(T), &full[s]);
                            if (len[i] > n_own)
                                bulk_g2s(dstw + (size_t)n_own * sizeof(T), gx + (lo[i] + n_own - g0), (len[i] - n_own) * (unsigned)sizeof(T), &full[s]);
                        }
                        at += len[i];
                    }
                if (++s == stages) { s = 0; parity ^= 1u; }
            }
        }
    } else {
        const unsigned lane = tid & 31u, cw = (tid >> 5) - 1u;             // consumer warp 0..14
        if constexpr (DIST) {
            const unsigned lane_id = tid - 32, n_lanes = kRingThreads - 32;
            halo_push<T>(halo, x, epoch, (uint64_t)blockIdx.x * n_lanes + lane_id, (uint64_t)gridDim.x * n_lanes);
            asm volatile("bar.sync 1, %0;" ::"r"(n_lanes) : "memory");
            if (lane_id == 0) halo_arrive(halo, epoch, gridDim.x);
        }
        unsigned s = 0, parity = 0;
        for (unsigned i = 0; i < n_my; ++i) {
            mbar_wait(&full[s], parity);
            const SellDesc d = s_desc[s];
            const unsigned char* base = smem_raw + (size_t)s * stage_bytes;
            const uint8_t* sc = base;
            const uint16_t* sp = reinterpret_cast<const uint16_t*>(base + o_cols);
            const uint8_t* sl = base + o_len;
            const unsigned* so = reinterpret_cast<const unsigned*>(base + o_soff);
            const T* sx = reinterpret_cast<const T*>(base + o_x);
            const T* dict = reinterpret_cast<const T*>(base + o_dict);
            for (unsigned q = cw; q < d.n_slices; q += kConsumerWarps) {
                const unsigned row = 32u * q + lane;
                const bool live = row < d.rows;
                const unsigned off = so[q], w = (so[q + 1] - off) >> 5;
                const unsigned len = live ? (unsigned)sl[row] : 0u;
                T wv = T(0);
                if constexpr (DOT) { if (live) wv = d.self_pos != ~0u ? sx[d.self_pos + row] : __ldg((const T*)dot.w + d.r0 + row); }
                T sum = T(0);
                const uint8_t* pc = sc + off + lane;
                const uint16_t* pp = sp + off + lane;
                // two entries per trip, all four index loads first, then the four operand loads; the trip count is the slice's
                // width (warp-uniform).  Positions are byte offsets into the staged windows.  When every row of the slice has the
                // slice's width (the common case) nothing is predicated; otherwise a lane's own length predicates the additions
                // (padding is loaded, never added).  Hand-shaped: the compiler's own unrolling of the simple loop cost 150
                // instructions per 7-entry slice.
                const unsigned char* sxb = reinterpret_cast<const unsigned char*>(sx);
                unsigned j = 0;
                if (__all_sync(0xffffffffu, len == w)) {
#pragma unroll 1
                    for (; j + 2 <= w; j += 2) {
                        const unsigned c0 = pc[0], p0 = pp[0], c1 = pc[32], p1 = pp[32];
                        pc += 64; pp += 64;
                        const T x0 = *reinterpret_cast<const T*>(sxb + p0), d0 = dict[c0];
                        const T x1 = *reinterpret_cast<const T*>(sxb + p1), d1 = dict[c1];
                        sum = add_rn(sum, mul_rn(x0, d0));
                        sum = add_rn(sum, mul_rn(x1, d1));
                    }
                    if (j < w) sum = add_rn(sum, mul_rn(*reinterpret_cast<const T*>(sxb + pp[0]), dict[pc[0]]));
                } else {
#pragma unroll 1
                    for (; j + 2 <= w; j += 2) {
                        const unsigned c0 = pc[0], p0 = pp[0], c1 = pc[32], p1 = pp[32];
                        pc += 64; pp += 64;
                        const T x0 = *reinterpret_cast<const T*>(sxb + p0), d0 = dict[c0];
                        const T x1 = *reinterpret_cast<const T*>(sxb + p1), d1 = dict[c1];
                        if (j < len) sum = add_rn(sum, mul_rn(x0, d0));
                        if (j + 1 < len) sum = add_rn(sum, mul_rn(x1, d1));
                    }
                    if (j < w) {
                        const unsigned c0 = pc[0], p0 = pp[0];
                        const T x0 = *reinterpret_cast<const T*>(sxb + p0), d0 = dict[c0];
                        if (j < len) sum = add_rn(sum, mul_rn(x0, d0));
                    }
                }
                if (live) {
                    y[d.r0 + row] = sum;
                    if constexpr (DOT) acc += (double)mul_rn(wv, sum);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == stages) { s = 0; parity ^= 1u; }
        }
    }
    if constexpr (DIST) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(halo.ctr + 1, 1u) == gridDim.x - 1) {
                halo.ctr[1] = 0u;
                *halo.epoch = epoch;
            }
        }
    }
    if constexpr (DOT) finish_dot<T, kRingThreads>(acc, dot);
}

}  // namespace smb
