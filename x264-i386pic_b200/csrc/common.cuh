// common.cuh -- device helpers shared by the sm_100a kernels of libx264cuda.so.
//
// Arithmetic notes (all integer, bit-exact with the reference's C functions):
//  * SAD uses VABSDIFF4.U8.ACC (__vsadu4): 4 pixels per instruction.
//  * SATD/SA8D/hadamard_ac use the reference's own "two 16-bit lanes in one 32-bit word" form
//    (common/pixel.c:242-259): x + (y<<16) as a true integer sum, plain 32-bit add/sub for the
//    butterflies, and abs2() at the end.  Using the same word-level arithmetic as pixel.c makes
//    the result identical by construction, including on checkasm's overflow vectors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define X264CU_PAD 32            // PADH == PADV == 32 (common/frame.h:32-33)
#define X264CU_COST_MAX (1<<28)  // COST_MAX (encoder/me.h)
#define X264CU_MAXB 16

typedef uint32_t u32;

__device__ __forceinline__ int iclip3( int v, int lo, int hi ) { return min( max( v, lo ), hi ); }
__device__ __forceinline__ int imedian( int a, int b, int c ) { return max( min( a, b ), min( max( a, b ), c ) ); }

// pixel.c:255-259 abs2(): per-lane absolute value of a true-sum packed word.  The reference's mask
// s = ((a>>15)&0x10001)*0xffff is 0xffff in each half whose bit 15 is set; PRMT's sign-replicate mode builds the
// same word in one instruction (bytes 0,1 <- sign of byte 1; bytes 2,3 <- sign of byte 3).
__device__ __forceinline__ u32 abs2( u32 a )
{
    u32 s;
    asm( "prmt.b32 %0, %1, 0, 0xbb99;" : "=r"( s ) : "r"( a ) );
    return ( a + s ) ^ s;
}

// pixel.c:242-251
#define HADAMARD4( d0, d1, d2, d3, s0, s1, s2, s3 ) { \
    u32 t0_ = (s0) + (s1), t1_ = (s0) - (s1), t2_ = (s2) + (s3), t3_ = (s2) - (s3); \
    d0 = t0_ + t2_; d2 = t0_ - t2_; d1 = t1_ + t3_; d3 = t1_ - t3_; }

// Load 8 consecutive bytes starting at an arbitrarily aligned address as two little-endian words.
// Reads the three aligned words that cover them (callers guarantee 4 bytes of slack past the row).
__device__ __forceinline__ uint2 load8u( const uint8_t *p )
{
    uintptr_t a = (uintptr_t)p;
    const u32 *q = (const u32 *)( a & ~(uintptr_t)3 );
    u32 sh = ( (u32)a & 3u ) * 8u;
    u32 w0 = __ldg( q ), w1 = __ldg( q + 1 ), w2 = __ldg( q + 2 );
    return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
}
// 4 consecutive bytes, arbitrary alignment
__device__ __forceinline__ u32 load4u( const uint8_t *p )
{
    uintptr_t a = (uintptr_t)p;
    const u32 *q = (const u32 *)( a & ~(uintptr_t)3 );
    u32 sh = ( (u32)a & 3u ) * 8u;
    return __funnelshift_r( __ldg( q ), __ldg( q + 1 ), sh );
}
// 16 consecutive bytes, arbitrary alignment
__device__ __forceinline__ uint4 load16u( const uint8_t *p )
{
    uintptr_t a = (uintptr_t)p;
    const u32 *q = (const u32 *)( a & ~(uintptr_t)3 );
    u32 sh = ( (u32)a & 3u ) * 8u;
    u32 w0 = __ldg( q ), w1 = __ldg( q + 1 ), w2 = __ldg( q + 2 ), w3 = __ldg( q + 3 ), w4 = __ldg( q + 4 );
    return make_uint4( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ),
                       __funnelshift_r( w2, w3, sh ), __funnelshift_r( w3, w4, sh ) );
}

// rounding byte average (a+b+1)>>1 on 4 packed pixels: mc.c:56 / FILTER of mc.c:494
// = (a|b) - (((a^b) & 0xfe..fe) >> 1) per byte; the explicit LOP3 keeps it at 4 instructions (__vavgu4 expands to 5)
__device__ __forceinline__ u32 avg4( u32 a, u32 b )
{
    u32 t;
    asm( "lop3.b32 %0, %1, %2, 0xfefefefe, 0x28;" : "=r"( t ) : "r"( a ), "r"( b ) );     // (a ^ b) & c
    return ( a | b ) - ( t >> 1 );
}

// SAD of 8 pixels held as two words each
__device__ __forceinline__ u32 sad8( uint2 a, uint2 b ) { return __vsadu4( a.x, b.x ) + __vsadu4( a.y, b.y ); }

// byte k (0..3) of a word / of a word pair
__device__ __forceinline__ u32 byte_of( u32 w, int k ) { return ( w >> ( 8 * k ) ) & 0xffu; }

// Packed differences of one 8-pixel row in the layout of x264_pixel_satd_8x4 (pixel.c:297-300):
// d[k] = (p1[k]-p2[k]) + ((p1[k+4]-p2[k+4]) << 16), k = 0..3, as true-sum words.
__device__ __forceinline__ void expand_row8( uint2 a, u32 e[4] )
{
    u32 t02 = __byte_perm( a.x, a.y, 0x6420 );     // p0 p2 p4 p6
    u32 t13 = __byte_perm( a.x, a.y, 0x7531 );     // p1 p3 p5 p7
    e[0] = __byte_perm( t02, 0, 0x4240 );          // p0 | p4<<16
    e[2] = __byte_perm( t02, 0, 0x4341 );          // p2 | p6<<16
    e[1] = __byte_perm( t13, 0, 0x4240 );          // p1 | p5<<16
    e[3] = __byte_perm( t13, 0, 0x4341 );          // p3 | p7<<16
}
__device__ __forceinline__ void diff_row8( uint2 a, uint2 b, u32 d[4] )
{
    u32 ea[4], eb[4];
    expand_row8( a, ea );
    expand_row8( b, eb );
#pragma unroll
    for( int k = 0; k < 4; k++ ) d[k] = ea[k] - eb[k];
}
