// hdlz_frame.cuh — container framing around the deflate body: zlib (RFC 1950, what the reference
// reads and writes: header bytes deflate.py:753-757, `di = 2` at STARTD :644, Adler-32 :788-814),
// raw deflate (RFC 1951 only) and gzip (RFC 1952).  Device helpers shared by the inflate kernels.
#pragma once

#include "hdlz_common.cuh"

namespace hdlz {

struct Frame {
    uint32_t body;     // byte offset of the first deflate block
    uint32_t trailer;  // bytes that must follow the byte-aligned end of the last block (0, 4 or 8)
    uint32_t status;   // HDLZ_OK, HDLZ_ST_TRUNCATED or HDLZ_ST_BAD_HEADER
};

// Parses the container header of one stream.  zlib: the two header bytes are skipped like the
// reference does (checked only with HDLZ_F_VERIFY_HEADER).  gzip: magic, CM and the reserved FLG
// bits are always checked (the member cannot be located otherwise); FEXTRA / FNAME / FCOMMENT /
// FHCRC are skipped.
__device__ inline Frame parse_frame(const uint8_t *src, uint32_t n_in, uint32_t flags)
{
    Frame f = {0u, 0u, HDLZ_OK};
    if (flags & HDLZ_F_RAW) return f;
    if (flags & HDLZ_F_GZIP) {
        f.trailer = 8;
        if (n_in < 18) { f.status = HDLZ_ST_TRUNCATED; return f; }
        const uint32_t flg = src[3];
        if (src[0] != 0x1Fu || src[1] != 0x8Bu || src[2] != 8u || (flg & 0xE0u)) { f.status = HDLZ_ST_BAD_HEADER; return f; }
        uint32_t p = 10;
        if (flg & 4u) {                                         // FEXTRA
            if (p + 2 > n_in) { f.status = HDLZ_ST_TRUNCATED; return f; }
            p += 2u + (src[p] | ((uint32_t)src[p + 1] << 8));
        }
        for (uint32_t bit = 8u; bit <= 16u; bit <<= 1) {        // FNAME, FCOMMENT: zero-terminated
            if (flg & bit) {
                while (p < n_in && src[p]) ++p;
                ++p;
            }
        }
        if (flg & 2u) p += 2;                                   // FHCRC
        if (p > n_in) { f.status = HDLZ_ST_TRUNCATED; return f; }
        f.body = p;
        return f;
    }
    f.body = 2;
    f.trailer = 4;
    if (n_in < 2) { f.status = HDLZ_ST_TRUNCATED; return f; }
    if (flags & HDLZ_F_VERIFY_HEADER) {
        const uint32_t cmf = src[0], flg = src[1];
        if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u || (flg & 0x20u)) f.status = HDLZ_ST_BAD_HEADER;
    }
    return f;
}

// CRC-32 (reflected 0xEDB88320), four bits per step.  `nib` = the 16-entry table crc32_nibble_entry builds.
__host__ __device__ inline uint32_t crc32_nibble_entry(uint32_t i)
{
    uint32_t c = i;
    for (int k = 0; k < 4; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
    return c;
}

__device__ __forceinline__ uint32_t crc32_byte(uint32_t crc, uint32_t b, const uint32_t *nib)
{
    crc ^= b;
    crc = (crc >> 4) ^ nib[crc & 15u];
    crc = (crc >> 4) ^ nib[crc & 15u];
    return crc;
}

// CRC-32 of p[0 .. n) by one thread (the optional gzip trailer check; not a hot path).
__device__ inline uint32_t crc32_bytes(const uint8_t *p, uint32_t n, const uint32_t *nib)
{
    uint32_t crc = 0xFFFFFFFFu;
    uint32_t i = 0;
    for (; i < n && ((reinterpret_cast<uintptr_t>(p) + i) & 3u); ++i) crc = crc32_byte(crc, p[i], nib);
    for (; i + 4 <= n; i += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);
        crc = crc32_byte(crc, w & 255u, nib);
        crc = crc32_byte(crc, (w >> 8) & 255u, nib);
        crc = crc32_byte(crc, (w >> 16) & 255u, nib);
        crc = crc32_byte(crc, w >> 24, nib);
    }
    for (; i < n; ++i) crc = crc32_byte(crc, p[i], nib);
    return ~crc;
}

__device__ __forceinline__ uint32_t load_le32(const uint8_t *p)
{
    return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

}  // namespace hdlz
