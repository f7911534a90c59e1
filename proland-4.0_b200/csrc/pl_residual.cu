/*
 * pl_residual.cu -- the ResidualProducer decode on the device: integer work,
 * bit-exact by construction.
 *
 * Reference: ResidualProducer::readTile (terrain/sources/proland/dem/
 * ResidualProducer.cpp:268-340) opens the tile's blob as an in-memory TIFF
 * (TIFFClientOpen over util/mfs), reads its single DEFLATE strip with
 * TIFFReadEncodedStrip (libtiff 3.x + zlib 1.x, third-party and absent from the
 * reference tree) and converts little-endian int16 -> float * scale, adding the
 * upsampled parent tile when composing the root levels.  File format:
 * src/terrain/doc/overview.txt:147-216, writer preprocess/terrain/
 * HeightMipmap.cpp:561-655.
 *
 * Here:
 *   host   : the IFD of every blob is parsed (tags 256/257 size, 258/277 two
 *            8-bit samples, 259 compression 1 / 8 / 32946, 273/279 the strip) and
 *            the strips are copied to the device in one transfer
 *   kernel1: inflate_kernel -- RFC 1950/1951 inflate, ONE WARP PER TILE.  Lane 0
 *            walks the bit stream with a 64-bit bit buffer fed by prefetched
 *            aligned words and table lookups (a 10-bit first-level table for
 *            literal/length codes and an 8-bit one for distances, built per
 *            dynamic block in shared memory by all 32 lanes; longer codes fall
 *            back to the canonical count/symbol walk).  Literals and LZ77 matches
 *            land in a 4 KB ring of recent output in shared memory; the warp
 *            flushes finished kilobytes to a dense scratch stream in HBM and
 *            serves the rare match that reaches behind the ring from there.
 *   kernel2: residual_store_kernel -- dense int16 -> the pool's pitched rows:
 *            raw for an I16 pool, (float) z * scale (+ the tile in add_slot) for an
 *            F32 pool, lower-left w x w corner of the slot exactly like the
 *            reference's 197-stride CPU slot.
 *
 * Any conforming inflater yields the same bytes; parity is pinned on the
 * reference's own fixture terrain4/DEM.dat (sha1 of every inflated tile,
 * tests/golden/dem_dat.json).
 */
#include <cstring>
#include <vector>

#include "pl_internal.h"

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kRing = 4096;      /* bytes of recent output per warp in shared memory (a power of two) */
constexpr int kFlush = 1024;     /* ... of which finished units of this size go to HBM */
constexpr int kLitBits = 10;     /* first-level table of the literal/length code */
constexpr int kDistBits = 8;     /* first-level table of the distance code */

enum {
    INF_OK = 0,
    INF_BAD_HEADER = 1,
    INF_BAD_BLOCK = 2,
    INF_BAD_CODE = 3,
    INF_OVERRUN_IN = 4,
    INF_OVERRUN_OUT = 5,
    INF_BAD_DISTANCE = 6,
    INF_SHORT = 7
};

struct WarpTables {
    unsigned short lit_lut[1 << kLitBits];   /* (symbol << 4) | code length, 0 = not in the table */
    unsigned short dist_lut[1 << kDistBits];
    unsigned short lit_count[16], lit_sym[288];   /* canonical decode (codes longer than the table) */
    unsigned short dist_count[16], dist_sym[32];
    unsigned char lens[320];
};

/* The bit reader (every lane of the warp carries an identical copy).  Input comes in ALIGNED 32-bit words, one word prefetched ahead of the bit
 * buffer so that the load latency overlaps the decoding of the 32 bits before it (byte loads on the
 * critical path were a third of the old kernel's time).  Consuming bits behind the end of the strip is
 * detected by position (br_overrun), not by what those bits are. */
struct BitReader {
    const unsigned char *src;    /* strip start */
    unsigned int end;            /* strip length in bytes */
    int org;                     /* strip offset of byte 0 of word 0 (src + org is 4-byte aligned; may be < 0) */
    unsigned int k;              /* index of the prefetched word `nxt` */
    unsigned int merged;         /* strip offset just past the bytes already merged into buf */
    unsigned int nxt;
    unsigned long long buf;
    int cnt;
};

/* word k of the stream: strip bytes [org + 4k, org + 4k + 4).  Words behind the end of the strip are
 * not fetched (zero); inside the last word the bytes behind the end are whatever follows the strip in the
 * packed blob buffer (the TIFF directory): a stream that consumes them has `overrun` set and is rejected */
__device__ __forceinline__ unsigned int br_word(const BitReader &b, unsigned int k)
{
    const int lo = b.org + 4 * (int) k;
    return lo < (int) b.end ? __ldg(reinterpret_cast<const unsigned int *>(b.src + lo)) : 0u;
}
/* start reading at strip offset t */
__device__ __forceinline__ void br_seek(BitReader &b, unsigned int t)
{
    const unsigned int a = (unsigned int) (reinterpret_cast<unsigned long long>(b.src + t) & 3ull);
    b.org = (int) t - (int) a;
    const unsigned int w0 = br_word(b, 0);
    b.buf = (unsigned long long) (w0 >> (8 * a));
    b.cnt = 32 - 8 * (int) a;
    b.merged = (unsigned int) (b.org + 4);
    b.k = 1;
    b.nxt = br_word(b, 1);
}
__device__ __forceinline__ void br_init(BitReader &b, const unsigned char *src, unsigned int n)
{
    b.src = src; b.end = n;
    br_seek(b, 0);
}
/* at least 33 valid bits afterwards */
__device__ __forceinline__ void br_fill(BitReader &b)
{
    if (b.cnt <= 32) {
        b.buf |= (unsigned long long) b.nxt << b.cnt;
        b.cnt += 32;
        b.merged += 4;
        b.k += 1;
        b.nxt = br_word(b, b.k);
    }
}
__device__ __forceinline__ unsigned int br_peek(const BitReader &b, int n) { return (unsigned int) (b.buf & ((1ull << n) - 1)); }
__device__ __forceinline__ void br_drop(BitReader &b, int n)
{
    b.buf >>= n;
    b.cnt -= n;
}
__device__ __forceinline__ unsigned int br_bits(BitReader &b, int n)
{
    if (b.cnt < n) br_fill(b);
    const unsigned int v = br_peek(b, n);
    br_drop(b, n);
    return v;
}
/* n bits the caller knows are buffered (no refill test) */
__device__ __forceinline__ unsigned int br_take(BitReader &b, int n)
{
    const unsigned int v = br_peek(b, n);
    br_drop(b, n);
    return v;
}
/* bits were consumed behind the end of the strip: bytes merged so far minus whole bytes still buffered
 * pass the end.  Checked at every flush and block end, not per symbol: a stream running on garbage stays
 * bounded by the output capacity and the code tables, and is rejected at the next check */
__device__ __forceinline__ bool br_overrun(const BitReader &b)
{
    return b.merged > b.end && b.merged - (unsigned int) (b.cnt >> 3) > b.end;
}
/* strip offset of the next unread byte (call at a byte boundary) */
__device__ __forceinline__ unsigned int br_byte_pos(const BitReader &b) { return b.merged - (unsigned int) (b.cnt >> 3); }

__device__ __forceinline__ unsigned int bitrev(unsigned int v, int n) { return __brev(v) >> (32 - n); }

/* Build count/symbol arrays and the first-level table of a canonical prefix code from
 * lens[0..n) (all 32 lanes; lane 0 does the short serial parts).  Returns false for an
 * over-subscribed code. */
__device__ bool build_code(const unsigned char *lens, int n, unsigned short *count, unsigned short *sym,
                           unsigned short *lut, int lut_bits, int lane)
{
    __shared__ unsigned short next_code_sh[kWarpsPerCta][16];
    __shared__ unsigned short offs_sh[kWarpsPerCta][16];
    unsigned short *next_code = next_code_sh[threadIdx.x >> 5];
    unsigned short *offs = offs_sh[threadIdx.x >> 5];
    bool ok = true;
    if (lane == 0) {
        for (int l = 0; l < 16; ++l) count[l] = 0;
        for (int s = 0; s < n; ++s) count[lens[s]]++;
        int left = 1;
        for (int l = 1; l < 16; ++l) {
            left <<= 1;
            left -= count[l];
            if (left < 0) ok = false;
        }
        unsigned int code = 0;
        offs[1] = 0;
        next_code[0] = 0;
        for (int l = 1; l < 16; ++l) {
            code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
            next_code[l] = (unsigned short) code;
            if (l < 15) offs[l + 1] = offs[l] + count[l];
        }
        /* symbols sorted by (length, value); next_code advanced in the same order */
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (l) sym[offs[l]++] = (unsigned short) s;
        }
    }
    ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
    for (int k = lane; k < (1 << lut_bits); k += 32) lut[k] = 0;
    __syncwarp();
    if (!ok) return false;
    /* canonical codes: symbol s of length l gets next_code[l] + (its rank among length-l symbols).
     * The ranks follow from the sorted sym[] array: entries of length l start at start(l). */
    if (lane == 0) {
        unsigned int start = 0;
        for (int l = 1; l < 16; ++l) { offs[l] = (unsigned short) start; start += count[l]; }
    }
    __syncwarp();
    for (int l = 1; l <= lut_bits; ++l) {
        const int c = count[l], base = offs[l];
        for (int r = lane; r < c; r += 32) {
            const unsigned int code = next_code[l] + r;
            const unsigned int rev = bitrev(code, l);
            const unsigned short e = (unsigned short) ((sym[base + r] << 4) | l);
            for (unsigned int pad = rev; pad < (1u << lut_bits); pad += 1u << l) lut[pad] = e;
        }
    }
    __syncwarp();
    return true;
}

/* canonical walk for codes longer than the first-level table (lane 0) */
__device__ int slow_decode(BitReader &b, const unsigned short *count, const unsigned short *sym)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int) br_bits(b, 1);
        const int c = count[l];
        if (code - c < first) return sym[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

__device__ __forceinline__ int decode_sym(BitReader &b, const unsigned short *lut, int lut_bits,
                                          const unsigned short *count, const unsigned short *sym)
{
    if (b.cnt < 32) br_fill(b);
    const unsigned short e = lut[br_peek(b, lut_bits)];
    if (e) {
        br_drop(b, e & 15);
        return e >> 4;
    }
    return slow_decode(b, count, sym);
}

__constant__ unsigned short kLenBase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
__constant__ unsigned char kLenExtra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
__constant__ unsigned short kDistBase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
__constant__ unsigned char kDistExtra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
__constant__ unsigned char kClOrder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

struct InflateJob {
    unsigned long long in_off;   /* strip start in the device blob buffer */
    unsigned int in_len;
    unsigned int out_len;        /* expected: w * w * 2 */
    unsigned int compression;    /* 1 = stored strip, else zlib stream */
};

/* ring[flushed .. flushed + n) -> dst (n, flushed multiples of 16; both sides 16-byte aligned) */
__device__ __forceinline__ void flush_ring(const unsigned char *ring, unsigned char *dst, unsigned int flushed, unsigned int n, int lane)
{
    const uint4 *src4 = reinterpret_cast<const uint4 *>(ring + (flushed & (kRing - 1)));
    uint4 *dst4 = reinterpret_cast<uint4 *>(dst + flushed);
    for (unsigned int q = lane; q < n / 16; q += 32) dst4[q] = src4[q];
}

/* One warp per tile, and ALL 32 LANES DECODE THE SAME STREAM IN LOCKSTEP: every lane carries the same bit
 * buffer, reads the same table entries (shared-memory broadcasts) and so knows every symbol, length and
 * distance without a shuffle.  The cost of a warp is its instruction count whatever the number of active
 * lanes, so the redundancy is free -- and it removes what made one-lane decoding slow (measured: 161
 * instructions per symbol): the divergence bookkeeping around every branch of a lane-0-only region, the
 * hand-over of matches to the other lanes, the byte-by-byte match loop.  Lane 0 alone writes literals
 * into a ring of the last kRing output bytes in shared memory; matches are copied by all lanes; finished
 * kilobytes leave for the dense stream in HBM as 16-byte stores; a match that reaches behind the ring
 * (rare) reads its source back from there. */
__global__ void __launch_bounds__(kWarpsPerCta * 32) inflate_kernel(int n, const InflateJob *jobs, const unsigned char *in,
                                                                   unsigned char *out, size_t out_stride, int *status)
{
    __shared__ WarpTables tables[kWarpsPerCta];
    __shared__ __align__(16) unsigned char rings[kWarpsPerCta][kRing];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int job = blockIdx.x * kWarpsPerCta + wid;
    if (job >= n) return;
    WarpTables &T = tables[wid];
    unsigned char *ring = rings[wid];
    const InflateJob J = jobs[job];
    const unsigned char *src = in + J.in_off;
    unsigned char *dst = out + (size_t) job * out_stride;
    const unsigned int cap = J.out_len;
    constexpr unsigned int M = kRing - 1;

    if (J.compression == 1) {   /* uncompressed strip */
        for (unsigned int k = lane; k < min(J.in_len, cap); k += 32) dst[k] = __ldg(src + k);
        if (lane == 0) status[job] = J.in_len == cap ? INF_OK : INF_SHORT;
        return;
    }

    BitReader b;
    br_init(b, src, J.in_len);
    int err = INF_OK;
    unsigned int pos = 0, flushed = 0;     /* bytes produced / bytes already in HBM (a multiple of kFlush) */
    {   /* zlib header: CM = 8, no preset dictionary, header checksum */
        const unsigned int cmf = br_bits(b, 8), flg = br_bits(b, 8);
        if ((cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32) || ((cmf << 8) | flg) % 31 != 0) err = INF_BAD_HEADER;
    }

    /* finished flush units of the ring -> HBM; the warp must have met since the last ring write */
    auto flush = [&]() {
        while (pos - flushed >= (unsigned int) kFlush) {
            flush_ring(ring, dst, flushed, kFlush, lane);
            flushed += kFlush;
        }
        __syncwarp();
    };

    int last = 0;
    while (!err && !last) {
        last = (int) br_bits(b, 1);
        const int type = (int) br_bits(b, 2);

        if (type == 0) {   /* stored block */
            br_drop(b, b.cnt & 7);                 /* to the byte boundary */
            const unsigned int len = br_bits(b, 16);
            const unsigned int nlen = br_bits(b, 16);
            if ((len ^ nlen) != 0xffffu) err = INF_BAD_BLOCK;
            const unsigned int src_pos = br_byte_pos(b);              /* next unread byte */
            if (!err && src_pos + len > b.end) err = INF_OVERRUN_IN;
            if (!err && pos + len > cap) err = INF_OVERRUN_OUT;
            if (err) break;
            /* through the ring, a flush unit at a time (later matches may refer to these bytes) */
            unsigned int done = 0;
            __syncwarp();
            while (done < len) {
                const unsigned int piece = min(len - done, (unsigned int) kFlush - (pos - flushed));
                for (unsigned int k = lane; k < piece; k += 32) ring[(pos + k) & M] = __ldg(src + src_pos + done + k);
                pos += piece;
                done += piece;
                __syncwarp();
                flush();
            }
            br_seek(b, src_pos + len);   /* restart the bit reader after the stored bytes */
            continue;
        }
        if (type == 3) { err = INF_BAD_BLOCK; break; }

        int nlen = 288, ndist = 30;
        __syncwarp();   /* nobody still reads the tables of the previous block */
        if (type == 1) {   /* fixed code */
            for (int s = lane; s < 288; s += 32) T.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
            for (int s = lane; s < 32; s += 32) T.lens[288 + s] = 5;
            ndist = 30;
        } else {           /* dynamic code: read the code lengths */
            nlen = (int) br_bits(b, 5) + 257;
            ndist = (int) br_bits(b, 5) + 1;
            const int ncode = (int) br_bits(b, 4) + 4;
            if (nlen > 286 || ndist > 30) { err = INF_BAD_BLOCK; break; }
            if (lane < 19) T.lens[lane] = 0;
            __syncwarp();
            for (int k = 0; k < ncode; ++k) {
                const unsigned int v = br_bits(b, 3);
                if (lane == 0) T.lens[kClOrder[k]] = (unsigned char) v;
            }
            __syncwarp();
            /* the code-length code reuses the literal tables (7-bit codes fit the first-level table) */
            if (!build_code(T.lens, 19, T.lit_count, T.lit_sym, T.lit_lut, kLitBits, lane)) { err = INF_BAD_CODE; break; }
            {
                /* the 19 code-length lengths are in use through the tables only, which are already
                 * built, so lens[] is overwritten (by lane 0; every lane tracks the previous length) */
                int idx = 0, prev_len = 0;
                while (idx < nlen + ndist && !err) {
                    const int s = decode_sym(b, T.lit_lut, kLitBits, T.lit_count, T.lit_sym);
                    if (s < 0) { err = INF_BAD_CODE; break; }
                    if (s < 16) {
                        if (lane == 0) T.lens[idx] = (unsigned char) s;
                        if (idx == 256 && s == 0) err = INF_BAD_CODE;   /* no end-of-block code */
                        idx++;
                        prev_len = s;
                    } else {
                        int prev = 0, rep;
                        if (s == 16) {
                            if (idx == 0) { err = INF_BAD_CODE; break; }
                            prev = prev_len;
                            rep = 3 + (int) br_bits(b, 2);
                        } else if (s == 17) {
                            rep = 3 + (int) br_bits(b, 3);
                        } else {
                            rep = 11 + (int) br_bits(b, 7);
                        }
                        if (idx + rep > nlen + ndist) { err = INF_BAD_CODE; break; }
                        if (idx <= 256 && idx + rep > 256 && prev == 0) err = INF_BAD_CODE;   /* lens[256] == 0 */
                        if (lane == 0)
                            for (int q = 0; q < rep; ++q) T.lens[idx + q] = (unsigned char) prev;
                        idx += rep;
                        prev_len = prev;
                    }
                }
            }
            if (err) break;
            __syncwarp();
        }
        /* distance lengths follow the literal/length lengths in lens[]; build both codes */
        __syncwarp();
        {
            const unsigned char *dl = T.lens + (type == 1 ? 288 : nlen);
            const bool okd = build_code(dl, ndist, T.dist_count, T.dist_sym, T.dist_lut, kDistBits, lane);
            const bool okl = build_code(T.lens, nlen, T.lit_count, T.lit_sym, T.lit_lut, kLitBits, lane);
            /* incomplete distance codes (a single distance code) are legal; over-subscription is not */
            if (!okd || !okl) { err = INF_BAD_CODE; break; }
        }

        /* symbols of the block.  Bit budget: decode_sym leaves >= 33 - 15 = 18 bits, enough for the length
         * extra bits (<= 5); one refill test before the distance code covers its 15 + 13 bits.  The output
         * bound is tested where it matters -- before a flush unit leaves for HBM and at the end: the ring
         * absorbs the at most kFlush + 258 bytes a corrupt stream can overshoot by in between. */
        /* Malformed input is noted in `bad` and acted upon at the next flush or block end instead of
         * branching out of the loop at every test (each early exit costs 4-5 instructions of reconvergence
         * bookkeeping per symbol): table indices are clamped, the ring index is masked, so decoding garbage
         * for at most one flush unit touches nothing outside the warp's own tables and ring. */
        int bad = 0;     /* bit 0: bad code, bit 1: distance reaches before the start of the output */
        for (;;) {
            const int sym = decode_sym(b, T.lit_lut, kLitBits, T.lit_count, T.lit_sym);
            if (sym < 256) {
                bad |= sym < 0 ? 1 : 0;
                if (lane == 0) ring[pos & M] = (unsigned char) sym;
                pos += 1;
            } else if (sym == 256) {
                break;
            } else {
                bad |= sym > 285 ? 1 : 0;
                const int li = min(sym - 257, 28);
                const unsigned int len = kLenBase[li] + br_take(b, kLenExtra[li]);
                int ds = decode_sym(b, T.dist_lut, kDistBits, T.dist_count, T.dist_sym);
                bad |= (ds < 0 || ds > 29) ? 1 : 0;
                ds = min(max(ds, 0), 29);
                const unsigned int dist = kDistBase[ds] + br_take(b, kDistExtra[ds]);
                bad |= dist > pos ? 2 : 0;
                __syncwarp();   /* lane 0's literals are in the ring */
                if (dist < (unsigned int) kRing) {
                    /* overlapping matches repeat their first `dist` bytes, all of which exist already */
                    if (dist >= len) {
                        for (unsigned int k = lane; k < len; k += 32) ring[(pos + k) & M] = ring[(pos - dist + k) & M];
                    } else {
                        for (unsigned int k = lane; k < len; k += 32) ring[(pos + k) & M] = ring[(pos - dist + (k % dist)) & M];
                    }
                } else if (!bad) {
                    /* the source lies at least kRing - 258 bytes behind pos: flushed long ago (dist > len) */
                    for (unsigned int k = lane; k < len; k += 32) ring[(pos + k) & M] = dst[pos - dist + k];
                }
                pos += len;
                __syncwarp();   /* the copy is complete before anybody writes behind it */
            }
            if (pos - flushed >= (unsigned int) kFlush) {
                if (bad) break;
                if (pos > cap) { err = INF_OVERRUN_OUT; break; }
                if (br_overrun(b)) { err = INF_OVERRUN_IN; break; }
                __syncwarp();
                flush();
            }
        }
        if (bad) err = (bad & 1) ? INF_BAD_CODE : INF_BAD_DISTANCE;
        if (!err && pos > cap) err = INF_OVERRUN_OUT;
        if (br_overrun(b)) err = INF_OVERRUN_IN;      /* also overrides what the garbage decoded to */
        __syncwarp();
        if (!err) flush();                            /* never past the tile's capacity */
    }
    if (!err && pos != cap) err = INF_SHORT;
    __syncwarp();
    if (!err) {   /* the tail: whole 16-byte words, then bytes */
        const unsigned int rest = pos - flushed, r16 = rest & ~15u;
        flush_ring(ring, dst, flushed, r16, lane);
        for (unsigned int k = r16 + lane; k < rest; k += 32) dst[flushed + k] = ring[(flushed + k) & M];
    }
    if (lane == 0) status[job] = err;
}

struct StoreJob {
    int width;       /* w of this tile */
    int out_slot;
    int add_slot;    /* -1: none */
    int channels;    /* byte tiles (pl_ortho_decode_batch): samples per texel in the dense stream */
};

/* dense int16 (w x w) -> pitched pool rows; ResidualProducer.cpp:321-338 */
template <bool F32>
__global__ void __launch_bounds__(256) residual_store_kernel(const StoreJob *jobs, const unsigned char *dense, size_t dense_stride,
                                                             unsigned char *pool, size_t slot_bytes, int pitch, float scale)
{
    const StoreJob J = jobs[blockIdx.x];
    const short *src = reinterpret_cast<const short *>(dense + (size_t) blockIdx.x * dense_stride);
    const int w = J.width;
    for (int k = threadIdx.x; k < w * w; k += blockDim.x) {
        const int j = k / w, i = k - j * w;
        const short z = src[k];
        if (F32) {
            float *dst = reinterpret_cast<float *>(pool + (size_t) J.out_slot * slot_bytes);
            const float zs = (float) z * scale;
            float v = zs;
            if (J.add_slot >= 0) {
                const float *add = reinterpret_cast<const float *>(pool + (size_t) J.add_slot * slot_bytes);
                v = add[(size_t) j * pitch + i] + zs;
            }
            dst[(size_t) j * pitch + i] = v;
        } else {
            short *dst = reinterpret_cast<short *>(pool + (size_t) J.out_slot * slot_bytes);
            dst[(size_t) j * pitch + i] = z;
        }
    }
}

/* dense bytes (w x w x channels, what TIFFReadEncodedStrip hands OrthoCPUProducer, OrthoCPUProducer.cpp:226-231)
 * -> RGBA8 texels of an ortho pool; channels the file does not have are written as 0 */
__global__ void __launch_bounds__(256) ortho_store_kernel(const StoreJob *jobs, const unsigned char *dense, size_t dense_stride,
                                                          unsigned char *pool, size_t slot_bytes)
{
    const StoreJob J = jobs[blockIdx.x];
    const unsigned char *src = dense + (size_t) blockIdx.x * dense_stride;
    uint32_t *dst = reinterpret_cast<uint32_t *>(pool + (size_t) J.out_slot * slot_bytes);
    const int w = J.width, ch = J.channels;
    for (int k = threadIdx.x; k < w * w; k += blockDim.x) {
        uint32_t t = 0;
        for (int c = 0; c < ch; ++c) t |= (uint32_t) src[(size_t) k * ch + c] << (8 * c);
        dst[k] = t;
    }
}

/* ResidualProducer::upsample (ResidualProducer.cpp:342-384): the (ts + 5)^2 tile of the next root
 * level from the quadrant (tx%2, ty%2) of its parent, in the CPU evaluation order of the reference
 * (NOT the GLSL mdot order): ((z1 + z2) * 9 - (z0 + z3)) / 16 on the axes, and for odd/odd texels the
 * running sum z += (f * g) * parent over dj = -1..2 (outer), di = -1..2 (inner).  Compiled with
 * --fmad=false: every product and sum rounds on its own, like the x86-64 SSE build of the reference. */
__global__ void __launch_bounds__(256) residual_upsample_kernel(const float *parent, float *result, int pitch, int ts, int px, int py)
{
    const int w = ts + 5;
    for (int k = threadIdx.x + blockIdx.x * blockDim.x; k < w * w; k += blockDim.x * gridDim.x) {
        const int j = k / w, i = k - j * w;
        const int cx = i / 2 + px, cy = j / 2 + py;
#define P(a, b) parent[(a) + (b) * pitch]
        float z;
        if (j % 2 == 0) {
            if (i % 2 == 0) {
                z = P(cx, cy);
            } else {
                const float z0 = P(cx - 1, cy), z1 = P(cx, cy), z2 = P(cx + 1, cy), z3 = P(cx + 2, cy);
                z = ((z1 + z2) * 9.0f - (z0 + z3)) / 16.0f;
            }
        } else {
            if (i % 2 == 0) {
                const float z0 = P(cx, cy - 1), z1 = P(cx, cy), z2 = P(cx, cy + 1), z3 = P(cx, cy + 2);
                z = ((z1 + z2) * 9.0f - (z0 + z3)) / 16.0f;
            } else {
                z = 0.0f;
                for (int dj = -1; dj <= 2; ++dj) {
                    const float f = (dj == -1 || dj == 2) ? -1 / 16.0f : 9 / 16.0f;
                    for (int di = -1; di <= 2; ++di) {
                        const float g = (di == -1 || di == 2) ? -1 / 16.0f : 9 / 16.0f;
                        z = z + (f * g) * P(cx + di, cy + dj);
                    }
                }
            }
        }
#undef P
        result[i + j * pitch] = z;
    }
}

inline unsigned int rd16(const uint8_t *p) { return (unsigned int) p[0] | ((unsigned int) p[1] << 8); }
inline unsigned int rd32(const uint8_t *p) { return rd16(p) | (rd16(p + 2) << 16); }

/* the first IFD of a little-endian baseline TIFF with one strip */
/* sample_bytes: bytes per texel the caller expects (2: the int16 residuals, stored as 2 x 8-bit or 1 x 16-bit
 * samples); 0: byte tiles with 1..4 samples of 8 bits (ColorMipmap::produceTile, preprocess/terrain/
 * ColorMipmap.cpp:312-325), *spp_out receives the sample count */
int parse_tiff(const uint8_t *blob, uint32_t size, int want_w, InflateJob *job, uint64_t base, int sample_bytes = 2,
               int *spp_out = nullptr)
{
    if (size < 8 || blob[0] != 'I' || blob[1] != 'I' || rd16(blob + 2) != 42) return -1;
    const uint32_t ifd = rd32(blob + 4);
    if ((uint64_t) ifd + 2 > size) return -2;
    const int n = (int) rd16(blob + ifd);
    if ((uint64_t) ifd + 2 + 12ull * n > size) return -3;
    uint32_t w = 0, h = 0, comp = 1, soff = 0, slen = 0, spp = 1, bps = 8, predictor = 1;
    for (int i = 0; i < n; ++i) {
        const uint8_t *e = blob + ifd + 2 + 12 * i;
        const unsigned int tag = rd16(e), type = rd16(e + 2), count = rd32(e + 4);
        const uint32_t val = type == 3 ? rd16(e + 8) : rd32(e + 8);
        switch (tag) {
        case 256: w = val; break;
        case 257: h = val; break;
        case 258:   /* BitsPerSample: two shorts fit the value field; more than two are an offset */
            if (count <= 2) {
                bps = rd16(e + 8);
                if (count == 2 && rd16(e + 10) != bps) bps = 0;
            } else {
                const uint32_t off = rd32(e + 8);
                if (type != 3 || count > 4 || (uint64_t) off + 2ull * count > size) return -8;
                bps = rd16(blob + off);
                for (unsigned int s = 1; s < count; ++s)
                    if (rd16(blob + off + 2 * s) != bps) bps = 0;
            }
            break;
        case 259: comp = val; break;
        case 273: soff = val; break;
        case 277: spp = val; break;
        case 279: slen = val; break;
        case 317: predictor = val; break;
        default: break;
        }
    }
    if ((int) w != want_w || h != w) return -4;
    if (predictor != 1) return -5;
    if (sample_bytes == 0) {
        if (bps != 8 || spp < 1 || spp > 4) return -5;
    } else if (spp * bps != 8u * (unsigned) sample_bytes) {
        return -5;
    }
    if (spp_out) *spp_out = (int) spp;
    if (comp != 1 && comp != 8 && comp != 32946) return -6;
    if ((uint64_t) soff + slen > size) return -7;
    job->in_off = base + soff;
    job->in_len = slen;
    job->out_len = w * w * (sample_bytes == 0 ? spp : (uint32_t) sample_bytes);
    job->compression = comp;
    return 0;
}

/* ---- the builder's side: HeightMipmap::computeResidual / encodeResidual / computeApproxTile
 * (preprocess/terrain/HeightMipmap.cpp:449-559), one CTA per tile ------------------------------- */

/* the height the parent approximation predicts for texel (i, j): the same taps and CPU evaluation order
 * as residual_upsample_kernel above (HeightMipmap.cpp:458-489 == ResidualProducer.cpp:349-380) */
__device__ __forceinline__ float hm_predict(const float *parent, int pitch, int i, int j, int px, int py)
{
    const int cx = i / 2 + px, cy = j / 2 + py;
#define P(a, b) parent[(a) + (b) * pitch]
    float z;
    if (j % 2 == 0) {
        if (i % 2 == 0) {
            z = P(cx, cy);
        } else {
            const float z0 = P(cx - 1, cy), z1 = P(cx, cy), z2 = P(cx + 1, cy), z3 = P(cx + 2, cy);
            z = ((z1 + z2) * 9.0f - (z0 + z3)) / 16.0f;
        }
    } else {
        if (i % 2 == 0) {
            const float z0 = P(cx, cy - 1), z1 = P(cx, cy), z2 = P(cx, cy + 1), z3 = P(cx, cy + 2);
            z = ((z1 + z2) * 9.0f - (z0 + z3)) / 16.0f;
        } else {
            z = 0.0f;
            for (int dj = -1; dj <= 2; ++dj) {
                const float f = (dj == -1 || dj == 2) ? -1 / 16.0f : 9 / 16.0f;
                for (int di = -1; di <= 2; ++di) {
                    const float g = (di == -1 || di == 2) ? -1 / 16.0f : 9 / 16.0f;
                    z = z + (f * g) * P(cx + di, cy + dj);
                }
            }
        }
    }
#undef P
    return z;
}

__global__ void __launch_bounds__(256) residual_encode_kernel(const pl_resid_enc_req *jobs, const unsigned char *hbase,
                                                              size_t hslot, int hpitch, unsigned char *abase, size_t aslot,
                                                              int apitch, unsigned char *rbase, size_t rslot, int rpitch,
                                                              float2 *stats)
{
    __shared__ float red_r[8], red_e[8];
    const pl_resid_enc_req J = jobs[blockIdx.x];
    const float *tile = reinterpret_cast<const float *>(hbase + (size_t) J.tile_slot * hslot);
    const float *parent = reinterpret_cast<const float *>(abase + (size_t) J.parent_slot * aslot);
    float *approx = reinterpret_cast<float *>(abase + (size_t) J.approx_slot * aslot);
    short *resid = reinterpret_cast<short *>(rbase + (size_t) J.resid_slot * rslot);
    const int w = J.tile_size + 5;
    const int px = 1 + (J.tx % 2) * J.tile_size / 2, py = 1 + (J.ty % 2) * J.tile_size / 2;
    float mr = 0.0f, me = 0.0f;
    for (int k = threadIdx.x; k < w * w; k += blockDim.x) {
        const int j = k / w, i = k - j * w;
        const float z = hm_predict(parent, apitch, i, j, px, py);
        const float t = tile[i + j * hpitch];
        const float diff = t - z;
        mr = fmaxf(fabsf(diff), mr);
        const short q = (short) (int) roundf(diff);          /* short(roundf(residual)): half away from zero */
        const float a = z + (float) q;
        me = fmaxf(fabsf(t - a), me);
        resid[i + j * rpitch] = q;
        approx[i + j * apitch] = a;
    }
    for (int s2 = 16; s2 > 0; s2 >>= 1) {
        mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, s2));
        me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, s2));
    }
    if ((threadIdx.x & 31) == 0) { red_r[threadIdx.x >> 5] = mr; red_e[threadIdx.x >> 5] = me; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) { mr = fmaxf(mr, red_r[q]); me = fmaxf(me, red_e[q]); }
        stats[blockIdx.x] = make_float2(mr, me);
    }
}

}  // namespace

static int decode_batch(pl_ctx *ctx, pl_pool *out, int n, const uint8_t *blobs, const uint64_t *offsets,
                        const uint32_t *sizes, const int32_t *widths, const int32_t *out_slots,
                        const int32_t *add_slots, float scale, int *channels_out)
{
    if (!ctx || !out || n < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (n == 0) return PL_OK;
    if (!blobs || !offsets || !sizes || !out_slots) return pl_set_error(PL_ERR_ARG, "NULL argument");
    const bool ortho = out->kind == PL_POOL_ORTHO_UN8x4 || out->kind == PL_POOL_NORM_UN8x4;
    if (!ortho && !widths) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (!ortho && out->kind != PL_POOL_RESID_F32 && out->kind != PL_POOL_RESID_I16)
        return pl_set_error(PL_ERR_ARG, "out is not a residual pool");
    if (add_slots && out->kind != PL_POOL_RESID_F32) return pl_set_error(PL_ERR_ARG, "add_slots needs an F32 pool");
    PL_CUDA(cudaSetDevice(ctx->device));

    /* pack the blobs back to back (8-byte aligned) and parse their IFDs */
    std::vector<InflateJob> jobs(n);
    std::vector<StoreJob> sjobs(n);
    std::vector<uint64_t> packed_off(n);
    uint64_t total = 0;
    int max_w = 0;
    for (int j = 0; j < n; ++j) {
        const int wj = ortho ? out->tile_w : widths[j];
        if (wj < 1 || wj > out->tile_w) return pl_set_error(PL_ERR_ARG, "tile %d: width %d exceeds the pool tile", j, wj);
        const int os = out_slots[j] == PL_SLOT_SCRATCH ? out->capacity : out_slots[j];
        const int as = !add_slots ? -1 : (add_slots[j] == PL_SLOT_SCRATCH ? out->capacity : add_slots[j]);
        if (os < 0 || os > out->capacity || as > out->capacity || (os == out->capacity && out->kind != PL_POOL_RESID_F32))
            return pl_set_error(PL_ERR_ARG, "tile %d: slot out of range", j);
        packed_off[j] = total;
        total += ((uint64_t) sizes[j] + 7) & ~7ull;
        int spp = 0;
        const int rc = parse_tiff(blobs + offsets[j], sizes[j], wj, &jobs[j], packed_off[j], ortho ? 0 : 2, &spp);
        if (rc) return pl_set_error(PL_ERR_CORRUPT, "tile %d: not a single-strip %s TIFF blob of width %d (code %d)", j,
                                    ortho ? "8-bit" : "16-bit", wj, rc);
        if (ortho && os >= out->capacity) return pl_set_error(PL_ERR_ARG, "tile %d: slot out of range", j);
        if (ortho && channels_out) {
            if (j > 0 && *channels_out != spp) return pl_set_error(PL_ERR_CORRUPT, "tile %d: %d channels, tile 0 has %d", j, spp, *channels_out);
            *channels_out = spp;
        }
        sjobs[j].width = wj;
        sjobs[j].out_slot = os;
        sjobs[j].add_slot = ortho || as < 0 ? -1 : as;
        sjobs[j].channels = spp;
        if (wj > max_w) max_w = wj;
    }
    const size_t dense_stride = ((size_t) max_w * max_w * (ortho ? 4 : 2) + 15) & ~(size_t) 15;
    const size_t bytes_in = (size_t) total + 16;

    /* staging: pinned host buffer -> device, one async copy */
    void *dev = nullptr;
    {
        const size_t job_off = (bytes_in + 15) & ~(size_t) 15;
        std::vector<uint8_t> stage(job_off + sizeof(InflateJob) * n + sizeof(StoreJob) * n, 0);
        for (int j = 0; j < n; ++j) memcpy(stage.data() + packed_off[j], blobs + offsets[j], sizes[j]);
        memcpy(stage.data() + job_off, jobs.data(), sizeof(InflateJob) * n);
        memcpy(stage.data() + job_off + sizeof(InflateJob) * n, sjobs.data(), sizeof(StoreJob) * n);
        int rc = pl_stage_requests(ctx, stage.data(), stage.size(), &dev);
        if (rc) return rc;
        /* scratch for the dense streams + status */
        const size_t scratch = dense_stride * n + sizeof(int) * n;
        if (scratch > ctx->resid_scratch_bytes) {
            PL_CUDA(cudaStreamSynchronize(ctx->stream));
            if (ctx->resid_scratch) cudaFree(ctx->resid_scratch);
            ctx->resid_scratch = nullptr;
            ctx->resid_scratch_bytes = 0;
            PL_CUDA(cudaMalloc(&ctx->resid_scratch, scratch + scratch / 4));
            ctx->resid_scratch_bytes = scratch + scratch / 4;
        }
        const unsigned char *d_in = static_cast<const unsigned char *>(dev);
        const InflateJob *d_jobs = reinterpret_cast<const InflateJob *>(d_in + job_off);
        const StoreJob *d_sjobs = reinterpret_cast<const StoreJob *>(d_in + job_off + sizeof(InflateJob) * n);
        unsigned char *d_dense = static_cast<unsigned char *>(ctx->resid_scratch);
        int *d_status = reinterpret_cast<int *>(d_dense + dense_stride * n);

        pl_timing_begin(ctx, PL_K_RESIDUAL, n);
        inflate_kernel<<<(n + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, 0, ctx->stream>>>(n, d_jobs, d_in, d_dense,
                                                                                                dense_stride, d_status);
        PL_CUDA(cudaGetLastError());
        if (ortho)
            ortho_store_kernel<<<n, 256, 0, ctx->stream>>>(d_sjobs, d_dense, dense_stride, out->base, out->slot_bytes);
        else if (out->kind == PL_POOL_RESID_F32)
            residual_store_kernel<true><<<n, 256, 0, ctx->stream>>>(d_sjobs, d_dense, dense_stride, out->base, out->slot_bytes, out->pitch, scale);
        else
            residual_store_kernel<false><<<n, 256, 0, ctx->stream>>>(d_sjobs, d_dense, dense_stride, out->base, out->slot_bytes, out->pitch, scale);
        PL_CUDA(cudaGetLastError());
        pl_timing_end(ctx);
        ctx->launches += 2;

        std::vector<int> status(n);
        PL_CUDA(cudaMemcpyAsync(status.data(), d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
        PL_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < n; ++j)
            if (status[j] != INF_OK)
                return pl_set_error(PL_ERR_CORRUPT, "tile %d: DEFLATE stream is corrupt (inflate code %d)", j, status[j]);
    }
    return PL_OK;
}

extern "C" int pl_residual_decode_batch(pl_ctx *ctx, pl_pool *out, int n, const uint8_t *blobs, const uint64_t *offsets,
                                        const uint32_t *sizes, const int32_t *widths, const int32_t *out_slots,
                                        const int32_t *add_slots, float scale)
{
    if (out && (out->kind == PL_POOL_ORTHO_UN8x4 || out->kind == PL_POOL_NORM_UN8x4))
        return pl_set_error(PL_ERR_ARG, "out is not a residual pool (byte tiles: pl_ortho_decode_batch)");
    return decode_batch(ctx, out, n, blobs, offsets, sizes, widths, out_slots, add_slots, scale, nullptr);
}

/* OrthoCPUProducer::doCreateTile, the TIFF branch (ortho/OrthoCPUProducer.cpp:205-232) */
extern "C" int pl_ortho_decode_batch(pl_ctx *ctx, pl_pool *out, int n, const uint8_t *blobs, const uint64_t *offsets,
                                     const uint32_t *sizes, const int32_t *out_slots, int *channels)
{
    if (out && out->kind != PL_POOL_ORTHO_UN8x4 && out->kind != PL_POOL_NORM_UN8x4)
        return pl_set_error(PL_ERR_ARG, "out is not an RGBA8 pool");
    int ch = 0;
    const int rc = decode_batch(ctx, out, n, blobs, offsets, sizes, nullptr, out_slots, nullptr, 1.0f, &ch);
    if (rc == PL_OK && channels && n > 0) *channels = ch;
    return rc;
}

extern "C" int pl_residual_upsample(pl_ctx *ctx, pl_pool *pool, int src_slot, int dst_slot, int tile_size, int tx, int ty)
{
    if (!ctx || !pool) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (pool->kind != PL_POOL_RESID_F32) return pl_set_error(PL_ERR_ARG, "pl_residual_upsample needs an F32 residual pool");
    const int src = src_slot == PL_SLOT_SCRATCH ? pool->capacity : src_slot;
    const int dst = dst_slot == PL_SLOT_SCRATCH ? pool->capacity : dst_slot;
    if (src < 0 || src > pool->capacity || dst < 0 || dst > pool->capacity || src == dst)
        return pl_set_error(PL_ERR_ARG, "pl_residual_upsample: slots %d -> %d", src_slot, dst_slot);
    /* parent reads span [px - 1, (ts + 4) / 2 + px + 2]: inside the parent tile for ts <= tile_w - 5 */
    if (tile_size < 2 || tile_size % 2 != 0 || tile_size + 5 > pool->tile_w || tx < 0 || ty < 0)
        return pl_set_error(PL_ERR_ARG, "pl_residual_upsample: tile size %d does not fit a %d pool", tile_size, pool->tile_w);
    PL_CUDA(cudaSetDevice(ctx->device));
    const int px = 1 + (tx % 2) * tile_size / 2, py = 1 + (ty % 2) * tile_size / 2;
    const float *parent = reinterpret_cast<const float *>(pool->base + (size_t) src * pool->slot_bytes);
    float *result = reinterpret_cast<float *>(pool->base + (size_t) dst * pool->slot_bytes);
    const int w = tile_size + 5;
    pl_timing_begin(ctx, PL_K_RESIDUAL, 1);
    residual_upsample_kernel<<<(w * w + 255) / 256, 256, 0, ctx->stream>>>(parent, result, pool->pitch, tile_size, px, py);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}

/* HeightMipmap::buildResiduals for n tiles of one level (preprocess/terrain/HeightMipmap.cpp:255-324):
 * residual = heights - upsample(parent approximation), rounded to int16; approximation = upsample +
 * rounded residual -- what ResidualProducer reconstructs from the file. */
extern "C" int pl_residual_encode_batch(pl_ctx *ctx, pl_pool *heights, pl_pool *approx, pl_pool *resid, int n,
                                        const pl_resid_enc_req *reqs, float *max_residual, float *max_err)
{
    if (!ctx || !heights || !approx || !resid || n < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (n == 0) return PL_OK;
    if (!reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    if (heights->kind != PL_POOL_RESID_F32 || approx->kind != PL_POOL_RESID_F32 || resid->kind != PL_POOL_RESID_I16)
        return pl_set_error(PL_ERR_ARG, "pl_residual_encode_batch needs F32 height and approximation pools and an I16 residual pool");
    for (int j = 0; j < n; ++j) {
        const pl_resid_enc_req &q = reqs[j];
        /* parent reads span [px - 1, (ts + 4) / 2 + px + 2] <= ts + 4: inside a parent tile of the same container */
        if (q.tile_size < 2 || q.tile_size % 2 != 0 || q.tile_size + 5 > heights->tile_w || q.tile_size + 5 > approx->tile_w ||
            q.tile_size + 5 > resid->tile_w || q.tx < 0 || q.ty < 0)
            return pl_set_error(PL_ERR_ARG, "request %d: tile size %d does not fit the pools", j, q.tile_size);
        if (q.tile_slot < 0 || q.tile_slot >= heights->capacity || q.parent_slot < 0 || q.parent_slot >= approx->capacity ||
            q.approx_slot < 0 || q.approx_slot >= approx->capacity || q.resid_slot < 0 || q.resid_slot >= resid->capacity ||
            q.approx_slot == q.parent_slot)
            return pl_set_error(PL_ERR_ARG, "request %d: slot out of range", j);
    }
    PL_CUDA(cudaSetDevice(ctx->device));
    void *dev = nullptr;
    const size_t req_bytes = (sizeof(pl_resid_enc_req) * (size_t) n + 15) & ~(size_t) 15;
    std::vector<uint8_t> stage(req_bytes + sizeof(float2) * (size_t) n, 0);
    memcpy(stage.data(), reqs, sizeof(pl_resid_enc_req) * (size_t) n);
    int rc = pl_stage_requests(ctx, stage.data(), stage.size(), &dev);
    if (rc) return rc;
    float2 *d_stats = reinterpret_cast<float2 *>(static_cast<unsigned char *>(dev) + req_bytes);
    pl_timing_begin(ctx, PL_K_RESIDUAL, n);
    residual_encode_kernel<<<n, 256, 0, ctx->stream>>>(static_cast<const pl_resid_enc_req *>(dev), heights->base, heights->slot_bytes,
                                                       heights->pitch, approx->base, approx->slot_bytes, approx->pitch, resid->base,
                                                       resid->slot_bytes, resid->pitch, d_stats);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (max_residual || max_err) {
        std::vector<float2> st(n);
        PL_CUDA(cudaMemcpyAsync(st.data(), d_stats, sizeof(float2) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
        PL_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < n; ++j) {
            if (max_residual) max_residual[j] = st[j].x;
            if (max_err) max_err[j] = st[j].y;
        }
    }
    return PL_OK;
}
