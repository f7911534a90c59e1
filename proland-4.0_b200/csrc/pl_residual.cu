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

/* ------------------------------------------------------------------------
 * DEFLATE on the device in two kernels.
 *
 * A DEFLATE stream is one serial chain: the position of every code depends on all codes before it.  The round-1
 * kernel gave a whole warp to one stream (32 lanes decoding redundantly, matches copied through a 4 KB ring in
 * shared memory): 104 warp instructions per symbol, issue-bound at 445 K tiles/s, and half of the matches of a
 * 197 x 197 int16 tile reach further back than any ring that fits (zlib's window is 32 KB), i.e. an L2 round trip
 * on the chain per match.  Splitting the warp into lane groups (one stream per 8 / 4 lanes) was measured SLOWER
 * (264 K / 214 K tiles/s): groups diverge at every literal/match branch and the warp runs them one after the other.
 * What the batch offers instead is thousands of independent streams, so:
 *
 *   inflate_tokens_kernel   ONE LANE PER STREAM does the serial part only -- Huffman decoding -- with uniform
 *       control flow (one symbol per lane and iteration, the match branch is predicated work, block headers fall on
 *       the same iteration for streams written with the same zlib settings).  A lane keeps its bit buffer in
 *       registers and its code tables in shared memory (2.5 KB: 64 streams per SM) and writes 32-bit TOKENS:
 *       up to three literal bytes, or (length, distance).  Nothing on its chain waits for HBM.
 *   lz_resolve_kernel       ONE WARP PER STREAM turns tokens into bytes, 32 tokens at a time: a warp scan gives
 *       every token its output offset; then, 32 output bytes per step, each lane finds the token its byte belongs
 *       to (binary search over the 32 offsets by SHFL) and fetches the byte the match points at.  The copies of a
 *       step are independent loads in flight together -- throughput, not a chain; a source inside the step itself
 *       (runs: distance < 32) is resolved by passing values between lanes.  The warp then converts its finished
 *       tile into the pool's layout (what residual_store_kernel did as a third launch).
 * ------------------------------------------------------------------------ */
constexpr int kTokLanes = 32;              /* streams per CTA of the tokenizer */

enum {
    INF_OK = 0,
    INF_BAD_HEADER = 1,
    INF_BAD_BLOCK = 2,
    INF_BAD_CODE = 3,
    INF_OVERRUN_IN = 4,
    INF_OVERRUN_OUT = 5,
    INF_BAD_DISTANCE = 6,
    INF_SHORT = 7,
    INF_BAD_CHECKSUM = 8
};

/* A canonical prefix code is decoded WITHOUT a look-up table: with v = the next 15 bits of the stream, first bit on
 * top, the code's length is the smallest l with v < lim[l] (canonical codes, left-justified, grow with their length)
 * and its symbol is sym[off[l] + (v >> (15 - l))].  lim[] and off[] of the literal/length and of the distance code
 * live in REGISTERS (60 of them: the kernel runs few warps per SM, registers are free), the comparison chain is
 * straight-line code every lane executes alike -- no table misses, no divergence, and a stream needs only its sorted
 * symbols in shared memory: 1 KB instead of the 2.6 KB of first-level tables, i.e. 2.5 x the streams per SM.
 * The trailing pad makes the stride an odd number of words, so that the 32 lanes of a warp touching the same field
 * of their own tables hit 32 different banks. */
struct LaneTables {
    unsigned short lit_sym[288];             /* symbols sorted by (code length, value) */
    unsigned short dist_sym[32];
    unsigned short lim[16];                  /* build_code -> registers; the code-length code is decoded from here */
    short off[16];
    unsigned short cnt[16];                  /* build_code scratch */
    unsigned int pad_;
};
constexpr int kLensBytes = 320;             /* code lengths of a block header: per stream, in global scratch (read and
                                             * written by the header phase only, a few hundred accesses per block) */
static_assert(sizeof(LaneTables) % 8 == 4, "odd word stride");

/* The bit reader of a stream.  Input comes in ALIGNED 32-bit words, one word prefetched ahead of the bit
 * buffer so that the load latency overlaps the decoding of the 32 bits before it.  Consuming bits behind the end
 * of the strip is detected by position (br_overrun), not by what those bits are. */
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
__device__ __forceinline__ unsigned int br_peek(const BitReader &b, int n) { return (unsigned int) b.buf & ((1u << n) - 1u); }   /* n <= 16 */
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
 * pass the end */
__device__ __forceinline__ bool br_overrun(const BitReader &b)
{
    return b.merged > b.end && b.merged - (unsigned int) (b.cnt >> 3) > b.end;
}
/* strip offset of the next unread byte (call at a byte boundary) */
__device__ __forceinline__ unsigned int br_byte_pos(const BitReader &b) { return b.merged - (unsigned int) (b.cnt >> 3); }
/* the 4-byte big-endian Adler-32 behind the last block (RFC 1950); false: the strip ends before it */
__device__ __forceinline__ bool br_trailer(BitReader &b, unsigned int *adler)
{
    br_drop(b, b.cnt & 7);
    const unsigned int p = br_byte_pos(b);
    if (p + 4 > b.end) return false;
    *adler = ((unsigned int) __ldg(b.src + p) << 24) | ((unsigned int) __ldg(b.src + p + 1) << 16) |
             ((unsigned int) __ldg(b.src + p + 2) << 8) | (unsigned int) __ldg(b.src + p + 3);
    return true;
}
/* The tables of a canonical prefix code from lens[0..n) (one lane, serial): symbols sorted by (length, value), and
 * lim[] / off[] as described at LaneTables.  Returns false for an over-subscribed code. */
/* zlib's rule (inftrees.c inflate_table), which is what the reference's libtiff decodes with: an over-subscribed code
 * is an error; an INCOMPLETE one too, except a literal/length or distance code whose longest (i.e. only) length is 1 bit,
 * or one with no symbols at all; the code-length code (`codes`) must be complete */
__device__ __forceinline__ bool code_is_acceptable(int left, const unsigned short *count, bool codes)
{
    if (left == 0) return true;
    int maxl = 0;
    for (int l = 1; l < 16; ++l)
        if (count[l]) maxl = l;
    return maxl == 0 || (!codes && maxl == 1);
}

__device__ bool build_code(const unsigned char *lens, int n, unsigned short *count, unsigned short *lim, short *off,
                           unsigned short *sym, bool codes = false)
{
    for (int l = 0; l < 16; ++l) count[l] = 0;
    for (int s = 0; s < n; ++s) count[lens[s]]++;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    if (!code_is_acceptable(left, count, codes)) return false;
    /* canonical codes in order of (length, value): code = first code of the length + rank */
    unsigned int code = 0, idx = 0;
    for (int l = 1; l < 16; ++l) {
        code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
        const unsigned int c = count[l];
        lim[l] = (unsigned short) min((code + c) << (15 - l), 0xffffu);
        off[l] = (short) ((int) idx - (int) code);
        idx += c;
    }
    /* the slot of a symbol is the running offset of its length: count[] becomes those offsets */
    unsigned int run = 0;
    for (int l = 1; l < 16; ++l) { const unsigned int c = count[l]; count[l] = (unsigned short) run; run += c; }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l) sym[count[l]++] = (unsigned short) s;
    }
    return true;
}

struct CodeRegs { unsigned int lim[16]; int dlt[16]; int off1; };   /* dlt[k] = off[k + 1] - off[k] */

__device__ __forceinline__ void load_code(CodeRegs &c, const unsigned short *lim, const short *off)
{
#pragma unroll
    for (int k = 1; k < 16; ++k) { c.lim[k] = lim[k]; c.dlt[k] = k < 15 ? (int) off[k + 1] - (int) off[k] : 0; }
    c.lim[0] = 0; c.dlt[0] = 0;
    c.off1 = off[1];
}

/* one symbol of a code held in registers; MAXL: the longest code the alphabet allows.  -1: not a code.
 * The 15 comparisons are independent of each other (a mask m_k = -1 where v >= lim[k]); the length and the symbol
 * offset are sums over the masks (off[l] = off[1] + sum of the differences below l), added as trees: the decode
 * is on the per-symbol dependency chain of a latency-bound kernel, so its DEPTH counts, not its instruction count */
template <int MAXL, int NSYM>
__device__ __forceinline__ int decode_sym(BitReader &b, const CodeRegs &c, const unsigned short *sym)
{
    if (b.cnt < 32) br_fill(b);
    const unsigned int v = __brev((unsigned int) b.buf) >> 17;   /* the next 15 bits, first bit on top */
    const unsigned int nv = ~v;                                  /* lim + nv = lim - v - 1: negative iff v >= lim */
    int m[16], t[16];
#pragma unroll
    for (int k = 1; k <= MAXL; ++k) {
        m[k] = (int) (c.lim[k] + nv) >> 31;
        t[k] = m[k] & c.dlt[k];
    }
    m[0] = 0; t[0] = 0;
#pragma unroll
    for (int k = MAXL + 1; k < 16; ++k) { m[k] = 0; t[k] = 0; }
    const int ms = ((m[0] + m[1]) + (m[2] + m[3])) + ((m[4] + m[5]) + (m[6] + m[7])) + (((m[8] + m[9]) + (m[10] + m[11])) + ((m[12] + m[13]) + (m[14] + m[15])));
    const int ts = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7])) + (((t[8] + t[9]) + (t[10] + t[11])) + ((t[12] + t[13]) + (t[14] + t[15])));
    int l = 1 - ms;
    const int o = c.off1 + ts;
    const bool bad = l > MAXL;
    l = bad ? MAXL : l;
    br_drop(b, l);
    const int s = sym[min(max(o + (int) (v >> (15 - l)), 0), NSYM - 1)];
    return bad ? -1 : s;
}
/* the code-length code of a block header (<= 7 bits), decoded from the arrays in shared memory */
__device__ __forceinline__ int decode_cl(BitReader &b, const unsigned short *lim, const short *off, const unsigned short *sym)
{
    if (b.cnt < 32) br_fill(b);
    const unsigned int v = __brev((unsigned int) b.buf) >> 17;
    int l = 1;
#pragma unroll
    for (int k = 1; k <= 7; ++k)
        if (v >= lim[k]) l = k + 1;
    if (l > 7) { br_drop(b, 7); return -1; }
    br_drop(b, l);
    return sym[min(max(off[l] + (int) (v >> (15 - l)), 0), 18)];
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

/* Tokens (32 bits): bits 31:30 = n > 0: n literal bytes in bits 7:0, 15:8, 23:16 (output order);
 *                   bits 31:30 = 0: a match, length - 3 in bits 7:0, distance - 1 in bits 29:15.
 * A stream of L output bytes has at most L / 2 + 2 tokens (the worst case alternates single literals with matches
 * of length 3). */
struct TokenInfo {
    unsigned int ntok;       /* tokens written */
    unsigned int out_len;    /* bytes they expand to */
    int status;
    int stored;              /* 1: uncompressed strip, no tokens: the resolver copies the strip */
    unsigned int adler;      /* the zlib trailer (Adler-32 of the output, RFC 1950): the resolver checks it, */
    int has_adler;           /* as inflate() does at the end of the stream; 0 for uncompressed strips */
};

/* ---- kernel 1: Huffman decoding, one lane per stream -------------------------------------------------- */
__global__ void __launch_bounds__(kTokLanes) inflate_tokens_kernel(int n, const InflateJob *jobs, const unsigned char *in,
                                                                    unsigned int *tokens, size_t tok_stride, TokenInfo *info,
                                                                    unsigned char *lens_scratch)
{
    extern __shared__ __align__(16) unsigned char tok_smem[];
    __shared__ unsigned int len_tab[32], dist_tab[32];   /* base | extra bits << 16 (constant memory would serialise over the lanes) */
    if (threadIdx.x < 29) len_tab[threadIdx.x] = kLenBase[threadIdx.x] | ((unsigned int) kLenExtra[threadIdx.x] << 16);
    if (threadIdx.x < 30) dist_tab[threadIdx.x] = kDistBase[threadIdx.x] | ((unsigned int) kDistExtra[threadIdx.x] << 16);
    __syncthreads();
    const int job = blockIdx.x * kTokLanes + threadIdx.x;
    const bool real = job < n;
    LaneTables &T = reinterpret_cast<LaneTables *>(tok_smem)[threadIdx.x];
    InflateJob J;
    J.in_off = 0; J.in_len = 0; J.out_len = 0; J.compression = 1;
    if (real) J = jobs[job];
    unsigned int *tok = tokens + (size_t) (real ? job : 0) * tok_stride;
    unsigned char *lens = lens_scratch + (size_t) (real ? job : 0) * kLensBytes;
    const unsigned int cap = J.out_len;
    const unsigned int tok_cap = (unsigned int) tok_stride;

    /* The lanes of a warp must stay CONVERGED: a warp instruction then advances 32 streams.  Left to itself the
     * compiler reconverges lanes that broke out of nested loops only behind those loops (measured: every lane ran
     * alone, 2 960 warp instructions per symbol step).  So the decoder is a state machine with one flat loop; every
     * iteration starts with a warp vote, which is also a reconvergence point: a lane either parses a block header
     * (streams written with the same zlib settings reach them on the same iteration), decodes ONE symbol, or idles. */
    enum { ST_HEADER = 0, ST_SYMBOLS = 1, ST_DONE = 2 };
    CodeRegs LC, DC;        /* the literal/length and the distance code of the current block */
#pragma unroll
    for (int k = 0; k < 16; ++k) { LC.lim[k] = DC.lim[k] = 0; LC.dlt[k] = DC.dlt[k] = 0; }   /* a header loads the real ones */
    LC.off1 = DC.off1 = 0;
    int state = ST_HEADER;
    int err = INF_OK;
    BitReader b;
    unsigned int pos = 0, ntok = 0;
    unsigned int lit_acc = 0, lit_n = 0;      /* pending literal token */
    int last = 0;
    if (!real || J.compression == 1) {
        state = ST_DONE;                       /* uncompressed strip: nothing to decode, the resolver copies it */
        b.src = in; b.end = 0; b.org = 0; b.k = 0; b.merged = 0; b.nxt = 0; b.buf = 0; b.cnt = 0;
    } else {
        br_init(b, in + J.in_off, J.in_len);
        /* zlib header: CM = 8, no preset dictionary, header checksum */
        const unsigned int cmf = br_bits(b, 8), flg = br_bits(b, 8);
        if ((cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32) || ((cmf << 8) | flg) % 31 != 0) { err = INF_BAD_HEADER; state = ST_DONE; }
    }

    while (__any_sync(0xffffffffu, state != ST_DONE)) {
        if (state == ST_HEADER) {
            /* ---- a block header: stored bytes, or the code tables of the block ---- */
            last = (int) br_bits(b, 1);
            const int type = (int) br_bits(b, 2);
            if (type == 0) {   /* stored block: its bytes become literal tokens */
                br_drop(b, b.cnt & 7);                 /* to the byte boundary */
                const unsigned int len = br_bits(b, 16);
                const unsigned int nlen = br_bits(b, 16);
                const unsigned int src_pos = br_byte_pos(b);              /* next unread byte */
                if ((len ^ nlen) != 0xffffu) err = INF_BAD_BLOCK;
                else if (src_pos + len > b.end) err = INF_OVERRUN_IN;
                else if (pos + len > cap || ntok + len / 3 + 2 > tok_cap) err = INF_OVERRUN_OUT;
                if (!err) {
                    for (unsigned int k = 0; k < len; ++k) {
                        lit_acc |= (unsigned int) __ldg(b.src + src_pos + k) << (8 * lit_n);
                        if (++lit_n == 3) {
                            tok[ntok++] = (3u << 30) | lit_acc;
                            lit_acc = 0; lit_n = 0;
                        }
                    }
                    pos += len;
                    br_seek(b, src_pos + len);   /* restart the bit reader after the stored bytes */
                    if (last) state = ST_DONE;
                }
            } else if (type == 3) {
                err = INF_BAD_BLOCK;
            } else {
                int nlen = 288, ndist = 30;
                if (type == 1) {   /* fixed code */
                    for (int s = 0; s < 288; ++s) lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
                    for (int s = 0; s < 32; ++s) lens[288 + s] = 5;
                } else {           /* dynamic code: read the code lengths */
                    nlen = (int) br_bits(b, 5) + 257;
                    ndist = (int) br_bits(b, 5) + 1;
                    const int ncode = (int) br_bits(b, 4) + 4;
                    if (nlen > 286 || ndist > 30) err = INF_BAD_BLOCK;
                    if (!err) {
                        for (int s = 0; s < 19; ++s) lens[s] = 0;
                        for (int k = 0; k < ncode; ++k) lens[kClOrder[k]] = (unsigned char) br_bits(b, 3);
                        /* the code-length code reuses the literal tables (its codes have at most 7 bits) */
                        if (!build_code(lens, 19, T.cnt, T.lim, T.off, T.lit_sym, true)) err = INF_BAD_CODE;
                    }
                    int idx = 0, prev_len = 0;
                    while (!err && idx < nlen + ndist) {
                        const int s = decode_cl(b, T.lim, T.off, T.lit_sym);
                        if (s < 0) {
                            err = INF_BAD_CODE;
                        } else if (s < 16) {
                            lens[idx] = (unsigned char) s;
                            if (idx == 256 && s == 0) err = INF_BAD_CODE;   /* no end-of-block code */
                            idx++;
                            prev_len = s;
                        } else {
                            int prev = 0, rep;
                            if (s == 16) {
                                if (idx == 0) err = INF_BAD_CODE;
                                prev = prev_len;
                                rep = 3 + (int) br_bits(b, 2);
                            } else if (s == 17) {
                                rep = 3 + (int) br_bits(b, 3);
                            } else {
                                rep = 11 + (int) br_bits(b, 7);
                            }
                            if (idx + rep > nlen + ndist) err = INF_BAD_CODE;
                            if (idx <= 256 && idx + rep > 256 && prev == 0) err = INF_BAD_CODE;   /* lens[256] == 0 */
                            if (!err)
                                for (int q = 0; q < rep; ++q) lens[idx + q] = (unsigned char) prev;
                            idx += rep;
                            prev_len = prev;
                        }
                    }
                }
                /* distance lengths follow the literal/length lengths in lens[]; incomplete distance codes (a single
                 * distance code) are legal, over-subscription is not */
                if (!err) {
                    /* the fixed distance code has 32 five-bit codes (30 and 31 never valid in data): complete */
                    if (!build_code(lens + (type == 1 ? 288 : nlen), type == 1 ? 32 : ndist, T.cnt, T.lim, T.off, T.dist_sym)) err = INF_BAD_CODE;
                    load_code(DC, T.lim, T.off);
                    if (!build_code(lens, nlen, T.cnt, T.lim, T.off, T.lit_sym)) err = INF_BAD_CODE;
                    load_code(LC, T.lim, T.off);
                }
                state = ST_SYMBOLS;
            }
            if (err) state = ST_DONE;
        } else if (state == ST_SYMBOLS) {
            /* ---- one symbol.  Bit budget: decode_sym leaves >= 33 - 15 = 18 bits, enough for the length extra
             * bits (<= 5); the distance decode refills for its 15 + 13 bits ---- */
            const int sym = decode_sym<15, 288>(b, LC, T.lit_sym);
            if (sym == 256) {
                if (br_overrun(b)) err = INF_OVERRUN_IN;
                state = last ? ST_DONE : ST_HEADER;
            } else if (sym < 0 || sym > 285) {
                err = INF_BAD_CODE;
            } else if (ntok + 2 > tok_cap) {
                err = INF_OVERRUN_OUT;
            } else if (sym < 256) {
                lit_acc |= (unsigned int) sym << (8 * lit_n);
                pos += 1;
                if (++lit_n == 3) {
                    tok[ntok++] = (3u << 30) | lit_acc;
                    lit_acc = 0; lit_n = 0;
                }
            } else {
                const unsigned int le = len_tab[sym - 257];
                const unsigned int len = (le & 0xffffu) + br_take(b, (int) (le >> 16));
                const int ds = decode_sym<15, 32>(b, DC, T.dist_sym);
                const unsigned int de = dist_tab[min(max(ds, 0), 29)];
                const unsigned int dist = (de & 0xffffu) + br_take(b, (int) (de >> 16));
                if (ds < 0 || ds > 29) err = INF_BAD_CODE;
                else if (dist > pos) err = INF_BAD_DISTANCE;
                else {
                    if (lit_n) {
                        tok[ntok++] = (lit_n << 30) | lit_acc;
                        lit_acc = 0; lit_n = 0;
                    }
                    tok[ntok++] = ((dist - 1u) << 15) | (len - 3u);
                    pos += len;
                }
            }
            if (pos > cap) err = INF_OVERRUN_OUT;
            if (err) state = ST_DONE;
        }
    }
    if (!real) return;
    TokenInfo ti;
    if (J.compression == 1) {
        ti.ntok = 0; ti.out_len = J.in_len; ti.stored = 1;
        ti.adler = 0; ti.has_adler = 0;
        ti.status = J.in_len == cap ? INF_OK : INF_SHORT;
    } else {
        if (!err && lit_n) {
            if (ntok >= tok_cap) err = INF_OVERRUN_OUT;
            else tok[ntok++] = (lit_n << 30) | lit_acc;
        }
        if (!err && pos != cap) err = INF_SHORT;
        ti.adler = 0; ti.has_adler = 1;
        if (!err && !br_trailer(b, &ti.adler)) err = INF_OVERRUN_IN;
        ti.ntok = ntok; ti.out_len = pos; ti.status = err; ti.stored = 0;
    }
    info[job] = ti;
}


/* ------------------------------------------------------------------------
 * Small batches: ONE WARP PER STREAM (the round-1 kernel).  The two-kernel decoder above needs tens of thousands of
 * streams to fill the chip (a lane per stream); below ~6 000 streams its tokenizer's fixed chain time (11 ms: 17 K
 * symbols of 1 300 cycles each) loses against 2.2 us per tile of this kernel.  Output: the dense stream + TokenInfo
 * with stored = 2, i.e. lz_resolve_kernel only moves the finished tile into the pool.
 * ------------------------------------------------------------------------ */
namespace warpinf {
constexpr int kWarpsPerCta = 4;
constexpr int kRing = 4096;      /* bytes of recent output per warp in shared memory (a power of two) */
constexpr int kFlush = 1024;     /* ... of which finished units of this size go to HBM */
constexpr int kLitBits = 10;     /* first-level table of the literal/length code */
constexpr int kDistBits = 8;     /* first-level table of the distance code */
__device__ __forceinline__ unsigned int bitrev(unsigned int v, int n) { return __brev(v) >> (32 - n); }

struct WarpTables {
    unsigned short lit_lut[1 << kLitBits];   /* (symbol << 4) | code length, 0 = not in the table */
    unsigned short dist_lut[1 << kDistBits];
    unsigned short lit_count[16], lit_sym[288];   /* canonical decode (codes longer than the table) */
    unsigned short dist_count[16], dist_sym[32];
    unsigned char lens[320];
};

/* Build count/symbol arrays and the first-level table of a canonical prefix code from
 * lens[0..n) (all 32 lanes; lane 0 does the short serial parts).  Returns false for an
 * over-subscribed code. */
__device__ bool build_code(const unsigned char *lens, int n, unsigned short *count, unsigned short *sym,
                           unsigned short *lut, int lut_bits, int lane, bool codes = false)
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
        if (ok && !code_is_acceptable(left, count, codes)) ok = false;
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
__global__ void __launch_bounds__(kWarpsPerCta * 32) inflate_warp_kernel(int n, const InflateJob *jobs, const unsigned char *in,
                                                                        unsigned char *out, size_t out_stride, TokenInfo *info)
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
        if (lane == 0) { TokenInfo ti; ti.ntok = 0; ti.out_len = cap; ti.status = J.in_len == cap ? INF_OK : INF_SHORT; ti.stored = 2; ti.adler = 0; ti.has_adler = 0; info[job] = ti; }
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
            if (!build_code(T.lens, 19, T.lit_count, T.lit_sym, T.lit_lut, kLitBits, lane, true)) { err = INF_BAD_CODE; break; }
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
            /* the fixed distance code has 32 five-bit codes (30 and 31 never valid in data): complete */
            const bool okd = build_code(dl, type == 1 ? 32 : ndist, T.dist_count, T.dist_sym, T.dist_lut, kDistBits, lane);
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
                if (dist <= (unsigned int) (kRing - 258)) {
                    /* source and destination are both inside the ring and share none of its bytes (a distance closer
                     * to the ring size would have one lane overwrite the ring byte another lane is about to read:
                     * those take the HBM branch below).  Overlapping matches repeat their first `dist` bytes, all of
                     * which exist already */
                    if (dist >= len) {
                        for (unsigned int k = lane; k < len; k += 32) ring[(pos + k) & M] = ring[(pos - dist + k) & M];
                    } else {
                        for (unsigned int k = lane; k < len; k += 32) ring[(pos + k) & M] = ring[(pos - dist + (k % dist)) & M];
                    }
                } else if (!bad) {
                    /* the source lies at least kRing - 2 * 258 bytes behind pos: flushed long ago (dist > len) */
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
    unsigned int adler = 0;
    if (!err && !br_trailer(b, &adler)) err = INF_OVERRUN_IN;
    if (lane == 0) { TokenInfo ti; ti.ntok = 0; ti.out_len = pos; ti.status = err; ti.stored = 2; ti.adler = adler; ti.has_adler = 1; info[job] = ti; }
}

}  // namespace warpinf

struct StoreJob {
    int width;       /* w of this tile */
    int out_slot;
    int add_slot;    /* -1: none */
    int channels;    /* byte tiles (pl_ortho_decode_batch): samples per texel in the dense stream */
};

/* what the resolver's warp does with its finished tile */
enum { STORE_I16 = 0, STORE_F32 = 1, STORE_ORTHO = 2 };

/* ---- kernel 2: tokens -> bytes, one warp per stream; then the tile into the pool's layout --------------- */
constexpr int kResolveWarps = 4;
constexpr int kWarpPathBelow = 6000;     /* batches smaller than this take the warp-per-stream decoder (see warpinf) */

template <int STORE>
__global__ void __launch_bounds__(kResolveWarps * 32, 16) lz_resolve_kernel(int n, const InflateJob *jobs, const unsigned char *in,
                                                                        const unsigned int *tokens, size_t tok_stride,
                                                                        const TokenInfo *info, unsigned char *dense, size_t dense_stride,
                                                                        int *status, const StoreJob *sjobs, unsigned char *pool,
                                                                        size_t slot_bytes, int pitch, float scale)
{
    const int lane = threadIdx.x & 31;
    const int job = blockIdx.x * kResolveWarps + (threadIdx.x >> 5);
    if (job >= n) return;
    const TokenInfo ti = info[job];
    unsigned char *dst = dense + (size_t) job * dense_stride;
    int err = ti.status;
    if (!err && ti.stored == 2) {
        /* the dense stream is complete already (inflate_warp_kernel) */
    } else if (!err && ti.stored) {
        const unsigned char *src = in + jobs[job].in_off;
        for (unsigned int k = lane; k < ti.out_len; k += 32) dst[k] = __ldg(src + k);
    } else if (!err) {
        const unsigned int *tok = tokens + (size_t) job * tok_stride;
        unsigned int pos = 0;       /* bytes finished */
        for (unsigned int t0 = 0; t0 < ti.ntok; t0 += 32) {
            const bool have = t0 + lane < ti.ntok;
            const unsigned int tw = have ? __ldg(tok + t0 + lane) : 0u;
            const unsigned int kind = tw >> 30;
            const unsigned int olen = !have ? 0u : (kind ? kind : (tw & 0xffu) + 3u);
            /* exclusive scan of the output lengths */
            unsigned int incl = olen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const unsigned int excl = incl - olen;
            const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
            for (unsigned int s = 0; s < total; s += 32) {
                const unsigned int q = s + lane;           /* output byte of this lane within the batch */
                /* the token that covers q: the last one whose offset is <= q (empty tokens sit at the end) */
                int t = 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const unsigned int oc = __shfl_sync(0xffffffffu, excl, t + d);
                    if (oc <= q) t += d;
                }
                const unsigned int tt = __shfl_sync(0xffffffffu, tw, t);
                const unsigned int te = __shfl_sync(0xffffffffu, excl, t);
                const bool active = q < total;
                const unsigned int r = q - te;            /* byte within the token */
                unsigned int val = 0;
                bool done = true;
                int from = 0;                              /* lane of this step that produces the source byte */
                if (active) {
                    if (tt >> 30) {
                        val = (tt >> (8 * r)) & 0xffu;
                    } else {
                        const unsigned int dist = ((tt >> 15) & 0x7fffu) + 1u;
                        if (dist > pos + q) {
                            err = INF_BAD_DISTANCE;        /* before the start of the output */
                        } else if (dist > lane) {
                            /* written by an earlier step (or batch): the stores of those are visible after the __syncwarp below */
                            val = __ldcg(dst + (pos + q - dist));
                        } else {
                            done = false;                  /* produced in this very step, by lane - dist */
                            from = lane - (int) dist;
                        }
                    }
                }
                /* runs: pass values down the lanes until every byte of the step is known (a source lane is always
                 * lower).  Pointer jumping: a lane whose source is itself still waiting adopts the source's source, so a
                 * chain of depth d (a run of distance 1 is 32 deep) resolves in log2 d + 1 rounds instead of d */
                while (__any_sync(0xffffffffu, !done)) {
                    const unsigned int v = __shfl_sync(0xffffffffu, val, from);
                    const int d = __shfl_sync(0xffffffffu, done ? 1 : 0, from);
                    const int f2 = __shfl_sync(0xffffffffu, from, from);
                    if (!done) {
                        if (d) { val = v; done = true; }
                        else from = f2;
                    }
                }
                if (active) dst[pos + q] = (unsigned char) val;
                __syncwarp();
            }
            pos += total;
        }
        if (!err && pos != ti.out_len) err = INF_SHORT;
    }
    err = (int) __reduce_max_sync(0xffffffffu, (unsigned int) err);   /* any lane's error is the stream's */
    if (err) {
        if (lane == 0) status[job] = err;
        return;
    }
    __syncwarp();

    /* the finished tile -> the pool (ResidualProducer.cpp:321-338; OrthoCPUProducer.cpp:226-231).  The same pass sums
     * the Adler-32 of the bytes it reads (s1 = 1 + sum b_i, s2 = len + sum (len - i) b_i, mod 65521): what inflate()
     * checks at the end of a zlib stream and TIFFReadEncodedStrip fails on.  A tile whose checksum is wrong is reported
     * corrupt; its slot holds the undefined bytes of a failed decode. */
    const StoreJob S = sjobs[job];
    const int w = S.width;
    const unsigned int len = ti.out_len;
    unsigned long long s1 = 0, s2 = 0;
    unsigned int covered = 0;                  /* bytes of the dense stream the store pass has read */
    if (STORE == STORE_ORTHO) {
        uint32_t *o = reinterpret_cast<uint32_t *>(pool + (size_t) S.out_slot * slot_bytes);
        const int ch = S.channels;
        for (int k = lane; k < w * w; k += 32) {
            uint32_t t = 0;
            for (int c = 0; c < ch; ++c) {
                const unsigned int i = (unsigned int) k * ch + c, v = __ldcg(dst + i);
                t |= v << (8 * c);
                s1 += v;
                s2 += (unsigned long long) (len - i) * v;
            }
            o[k] = t;
        }
        covered = (unsigned int) (w * w * ch);
    } else {
        const unsigned short *src = reinterpret_cast<const unsigned short *>(dst);
        int j = lane / w, i = lane - j * w;          /* texel k = i + j w, advanced without a division per texel */
        for (int k = lane; k < w * w; k += 32) {
            const unsigned int zu = __ldcg(src + k);
            const short z = (short) zu;
            const unsigned int lo = zu & 255u, hi = zu >> 8;
            s1 += lo + hi;
            s2 += (unsigned long long) (len - 2u * (unsigned int) k) * (lo + hi) - hi;
            if (STORE == STORE_F32) {
                float *o = reinterpret_cast<float *>(pool + (size_t) S.out_slot * slot_bytes);
                const float zs = (float) z * scale;
                float v = zs;
                if (S.add_slot >= 0) {
                    const float *add = reinterpret_cast<const float *>(pool + (size_t) S.add_slot * slot_bytes);
                    v = add[(size_t) j * pitch + i] + zs;
                }
                o[(size_t) j * pitch + i] = v;
            } else {
                short *o = reinterpret_cast<short *>(pool + (size_t) S.out_slot * slot_bytes);
                o[(size_t) j * pitch + i] = z;
            }
            i += 32;
            while (i >= w) { i -= w; ++j; }
        }
        covered = 2u * (unsigned int) (w * w);
    }
    if (ti.has_adler) {
        for (unsigned int i = covered + lane; i < len; i += 32) {      /* a stream longer than the tile (not the case for valid files) */
            const unsigned int v = __ldcg(dst + i);
            s1 += v;
            s2 += (unsigned long long) (len - i) * v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        }
        const unsigned int adler = (unsigned int) (((s2 + len) % 65521ull) << 16) | (unsigned int) ((s1 + 1ull) % 65521ull);
        if (covered > len || adler != ti.adler) err = INF_BAD_CHECKSUM;
    }
    if (lane == 0) status[job] = err;
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
    const bool root = J.parent_slot < 0;      /* level 0: stored as short(roundf(h)), its own approximation (produceTile :567-578) */
    const float *parent = reinterpret_cast<const float *>(abase + (size_t) (root ? 0 : J.parent_slot) * aslot);
    float *approx = reinterpret_cast<float *>(abase + (size_t) J.approx_slot * aslot);
    short *resid = reinterpret_cast<short *>(rbase + (size_t) J.resid_slot * rslot);
    const int w = J.tile_size + 5;
    const int px = 1 + (J.tx % 2) * J.tile_size / 2, py = 1 + (J.ty % 2) * J.tile_size / 2;
    float mr = 0.0f, me = 0.0f;
    for (int k = threadIdx.x; k < w * w; k += blockDim.x) {
        const int j = k / w, i = k - j * w;
        const float z = root ? 0.0f : hm_predict(parent, apitch, i, j, px, py);
        const float t = tile[i + j * hpitch];
        const float diff = t - z;
        mr = fmaxf(fabsf(diff), mr);
        const short q = (short) (int) roundf(diff);          /* short(roundf(residual)): half away from zero */
        const float a = root ? t : z + (float) q;           /* getApproxTile(0) is the height tile itself (:420-431) */
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

/* A residual / ortho archive resident in device memory (pl_blobs_create): the reference maps its .dat files into
 * the address space once (ResidualProducer.cpp:70-129, util/mfs) and reads tiles from there; here the whole file is
 * uploaded once and the decoders read the compressed strips in place -- no per-batch packing, no PCIe traffic. */
struct pl_blobs {
    pl_ctx *ctx;
    unsigned char *dev;
    uint64_t size;
};

static int decode_batch(pl_ctx *ctx, pl_pool *out, int n, const uint8_t *blobs, const uint64_t *offsets,
                        const uint32_t *sizes, const int32_t *widths, const int32_t *out_slots,
                        const int32_t *add_slots, float scale, int *channels_out, const pl_blobs *store = nullptr)
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
        if (store && offsets[j] + sizes[j] > store->size) return pl_set_error(PL_ERR_ARG, "tile %d: blob outside the archive", j);
        packed_off[j] = store ? offsets[j] : total;     /* resident archive: the strip is read where it lies */
        if (!store) total += ((uint64_t) sizes[j] + 7) & ~7ull;
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
        if (!store)
            for (int j = 0; j < n; ++j) memcpy(stage.data() + packed_off[j], blobs + offsets[j], sizes[j]);
        memcpy(stage.data() + job_off, jobs.data(), sizeof(InflateJob) * n);
        memcpy(stage.data() + job_off + sizeof(InflateJob) * n, sjobs.data(), sizeof(StoreJob) * n);
        int rc = pl_stage_requests(ctx, stage.data(), stage.size(), &dev);
        if (rc) return rc;
        /* scratch: the dense streams, status words, token streams and their counts */
        size_t max_out = 0;
        for (int j = 0; j < n; ++j) max_out = jobs[j].out_len > max_out ? jobs[j].out_len : max_out;
        const size_t tok_stride = (max_out / 2 + 4 + 3) & ~(size_t) 3;                  /* tokens per stream (see TokenInfo) */
        const size_t off_status = dense_stride * n;
        const size_t off_info = (off_status + sizeof(int) * n + 15) & ~(size_t) 15;
        const size_t off_tok = (off_info + sizeof(TokenInfo) * n + 15) & ~(size_t) 15;
        const size_t off_lens = off_tok + tok_stride * sizeof(unsigned int) * n;
        const size_t scratch = off_lens + (size_t) kLensBytes * n;
        if (scratch > ctx->resid_scratch_bytes) {
            PL_CUDA(cudaStreamSynchronize(ctx->stream));
            if (ctx->resid_scratch) cudaFree(ctx->resid_scratch);
            ctx->resid_scratch = nullptr;
            ctx->resid_scratch_bytes = 0;
            PL_CUDA(cudaMalloc(&ctx->resid_scratch, scratch + scratch / 4));
            ctx->resid_scratch_bytes = scratch + scratch / 4;
        }
        const unsigned char *d_stage = static_cast<const unsigned char *>(dev);
        const unsigned char *d_in = store ? store->dev : d_stage;
        const InflateJob *d_jobs = reinterpret_cast<const InflateJob *>(d_stage + job_off);
        const StoreJob *d_sjobs = reinterpret_cast<const StoreJob *>(d_stage + job_off + sizeof(InflateJob) * n);
        unsigned char *d_dense = static_cast<unsigned char *>(ctx->resid_scratch);
        int *d_status = reinterpret_cast<int *>(d_dense + off_status);
        TokenInfo *d_info = reinterpret_cast<TokenInfo *>(d_dense + off_info);
        unsigned int *d_tok = reinterpret_cast<unsigned int *>(d_dense + off_tok);
        unsigned char *d_lens = d_dense + off_lens;

        const size_t tsmem = sizeof(LaneTables) * kTokLanes;
        static bool attr_set = false;
        if (!attr_set) {
            PL_CUDA(cudaFuncSetAttribute(inflate_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tsmem));
            attr_set = true;
        }
        pl_timing_begin(ctx, PL_K_RESIDUAL, n);
        if (ctx->inflate_path == 1 || (ctx->inflate_path == 0 && n < kWarpPathBelow))
            warpinf::inflate_warp_kernel<<<(n + warpinf::kWarpsPerCta - 1) / warpinf::kWarpsPerCta, warpinf::kWarpsPerCta * 32, 0, ctx->stream>>>(
                n, d_jobs, d_in, d_dense, dense_stride, d_info);
        else
            inflate_tokens_kernel<<<(n + kTokLanes - 1) / kTokLanes, kTokLanes, tsmem, ctx->stream>>>(n, d_jobs, d_in, d_tok, tok_stride, d_info, d_lens);
        PL_CUDA(cudaGetLastError());
        const int rgrid = (n + kResolveWarps - 1) / kResolveWarps;
        if (ortho)
            lz_resolve_kernel<STORE_ORTHO><<<rgrid, kResolveWarps * 32, 0, ctx->stream>>>(n, d_jobs, d_in, d_tok, tok_stride, d_info, d_dense, dense_stride,
                                                                                      d_status, d_sjobs, out->base, out->slot_bytes, out->pitch, scale);
        else if (out->kind == PL_POOL_RESID_F32)
            lz_resolve_kernel<STORE_F32><<<rgrid, kResolveWarps * 32, 0, ctx->stream>>>(n, d_jobs, d_in, d_tok, tok_stride, d_info, d_dense, dense_stride,
                                                                                    d_status, d_sjobs, out->base, out->slot_bytes, out->pitch, scale);
        else
            lz_resolve_kernel<STORE_I16><<<rgrid, kResolveWarps * 32, 0, ctx->stream>>>(n, d_jobs, d_in, d_tok, tok_stride, d_info, d_dense, dense_stride,
                                                                                    d_status, d_sjobs, out->base, out->slot_bytes, out->pitch, scale);
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

extern "C" int pl_blobs_create(pl_ctx *ctx, const uint8_t *bytes, uint64_t size, pl_blobs **out)
{
    if (!ctx || !bytes || !out || size == 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    pl_blobs *b = new pl_blobs();
    b->ctx = ctx;
    b->size = size;
    b->dev = nullptr;
    /* 16 spare bytes: the bit reader fetches aligned words, the last of which may straddle the end */
    cudaError_t e = cudaMalloc(&b->dev, size + 16);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->dev + size, 0, 16, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->dev, bytes, size, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        if (b->dev) cudaFree(b->dev);
        delete b;
        return pl_set_error(PL_ERR_CUDA, "pl_blobs_create: %s", cudaGetErrorString(e));
    }
    *out = b;
    return PL_OK;
}

extern "C" void pl_blobs_destroy(pl_blobs *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->dev);
    delete b;
}

extern "C" int pl_residual_decode_stored(pl_ctx *ctx, pl_pool *out, const pl_blobs *store, const uint8_t *host_bytes, int n,
                                         const uint64_t *offsets, const uint32_t *sizes, const int32_t *widths,
                                         const int32_t *out_slots, const int32_t *add_slots, float scale)
{
    if (!store || store->ctx != ctx) return pl_set_error(PL_ERR_ARG, "the archive belongs to another context");
    if (out && (out->kind == PL_POOL_ORTHO_UN8x4 || out->kind == PL_POOL_NORM_UN8x4))
        return pl_set_error(PL_ERR_ARG, "out is not a residual pool");
    return decode_batch(ctx, out, n, host_bytes, offsets, sizes, widths, out_slots, add_slots, scale, nullptr, store);
}

extern "C" int pl_debug_inflate_path(pl_ctx *ctx, int path)
{
    if (!ctx || path < 0 || path > 2) return pl_set_error(PL_ERR_ARG, "bad argument");
    ctx->inflate_path = path;
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
        if (q.tile_slot < 0 || q.tile_slot >= heights->capacity || q.parent_slot < -1 || q.parent_slot >= approx->capacity ||
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
