// la_block.cuh -- per-block cost machinery shared by the search and decision kernels (search.cu, final.cu):
// row-lane and whole-block SAD/SATD, the qpel fetch of get_ref, and the two "backends" that say how a block's pixels
// are spread over lanes.  Everything here is inline device code; each kernel file is its own translation unit so that
// the code generation of one kernel family cannot perturb the other's (r01: register allocation of the decision
// kernel changed with unrelated edits of the search kernel while both lived in one file).
#pragma once
#include "lookahead.cuh"

// ---- tuning knobs, all measured on B200 (profiles/r01/tuning.md).  The thread-per-block kernels are sensitive to
// ptxas's choices under register pressure (if-conversion of the second-plane fetch, rematerialised row addresses), so
// the knobs are per kernel family and each family is its own translation unit.
// CTA tiles in blocks (lookahead.cuh: X264CU_S_TILE_* / X264CU_F_TILE_*): one block row per warp for the search kernels,
// 16x8 for the decision kernels (squarer tile = less halo in the L2->L1 fill traffic that bounds them)
#define S_TILE_W X264CU_S_TILE_W
#define S_TILE_H X264CU_S_TILE_H
#define F_TILE_W X264CU_F_TILE_W
#define F_TILE_H X264CU_F_TILE_H
// row loads as three 32-bit words (0), two aligned 64-bit words (1), or one 64-bit + one 32-bit word (2): RowSrc
#ifndef X264CU_SEARCH_LD64
#define X264CU_SEARCH_LD64 2
#endif
#ifndef X264CU_FINAL_LD64
#define X264CU_FINAL_LD64 2
#endif
// second-plane fetch of get_ref as PTX-predicated loads (1) or a C++ branch (0)
#ifndef X264CU_SEARCH_PTXPRED
#define X264CU_SEARCH_PTXPRED 1
#endif
#ifndef X264CU_FINAL_PTXPRED
#define X264CU_FINAL_PTXPRED 1
#endif
// __launch_bounds__ minimum CTAs per SM.  Search kernel: 32x8-block CTAs of 256 threads, 2 per SM (126 registers, 16 warps); with
// 32x4-block CTAs 3 / 4 / 5 / 6 per SM (118 / 116 / 96 / 80 registers) all run one context within 3 % of each other, but under the
// benchmarked load (4 contexts whose search and decision launches overlap) the 256-thread CTAs give 2 % more frames/s.
// Decision kernel: 4 CTAs of 128 threads at 128 registers (see X264CU_BIDIR_AHEAD).
// dedicated half-pel diamond for even centres in the thread-per-block search (1) or the generic qpel fetch only (0)
#ifndef X264CU_HPEL_EVEN
#define X264CU_HPEL_EVEN 1
#endif
#ifndef X264CU_MINB_SEARCH
#define X264CU_MINB_SEARCH 2
#endif
#ifndef X264CU_MINB_FINAL
#define X264CU_MINB_FINAL 4
#endif

#ifndef X264CU_ROWIF_ZERO
#define X264CU_ROWIF_ZERO 0
#endif
// Decision kernel: how much of a bidir candidate is requested before its transform starts -- 0: row by row (r01, r02a);
// 3: both motion-compensated predictions (8 rows x 2 references x up to 2 planes = up to 64 loads in flight per thread);
// 4: also the two zero-MV windows.  The kernel is bound by the latency of its L2 -> L1 stream (65 % long-scoreboard stalls);
// with row-by-row fetches a warp had 2 - 8 loads in flight.  Needs registers: 4 CTAs per SM at 128 (profiles/r02/tuning.md).
#ifndef X264CU_BIDIR_AHEAD
#define X264CU_BIDIR_AHEAD 4
#endif
#define GROUPS_PER_CTA 32
#define CTA_THREADS (GROUPS_PER_CTA*8)

static __constant__ uint8_t c_hpel_ref0[16] = {0,1,1,1,0,1,1,1,2,3,3,3,0,1,1,1};   // tables.c:183
static __constant__ uint8_t c_hpel_ref1[16] = {0,0,1,0,2,2,3,2,2,2,3,2,2,2,3,2};   // tables.c:184
// the same two tables as immediates (2 bits per entry): the index differs from lane to lane in the thread-per-block kernels,
// where an indexed constant load is replayed once per distinct index
__device__ __forceinline__ int hpel_ref0( int qidx ) { return (int)( ( 0x54fe5454u >> ( 2*qidx ) ) & 3u ); }
__device__ __forceinline__ int hpel_ref1( int qidx ) { return (int)( ( 0xbababa10u >> ( 2*qidx ) ) & 3u ); }

__device__ __forceinline__ int weight_px( int v, int scale, int denom, int offset )
{   // mc.c:117-118
    int r = denom >= 1 ? ( ( v*scale + ( 1 << ( denom-1 ) ) ) >> denom ) + offset : v*scale + offset;
    return min( max( r, 0 ), 255 );
}
__device__ __forceinline__ u32 weight4( u32 w, int scale, int denom, int offset )
{
    return (u32)weight_px( w & 0xff, scale, denom, offset ) | ( (u32)weight_px( ( w >> 8 ) & 0xff, scale, denom, offset ) << 8 )
         | ( (u32)weight_px( ( w >> 16 ) & 0xff, scale, denom, offset ) << 16 ) | ( (u32)weight_px( w >> 24, scale, denom, offset ) << 24 );
}

// ------------------------------------------------------------------------------------------------
// row-lane metrics
// ------------------------------------------------------------------------------------------------
struct Grp
{
    unsigned mask;    // the 8 lanes of this block
    int r;            // row owned by this lane
};

__device__ __forceinline__ int red8( int v, const Grp &g )
{
    v += __shfl_xor_sync( g.mask, v, 1 );
    v += __shfl_xor_sync( g.mask, v, 2 );
    v += __shfl_xor_sync( g.mask, v, 4 );
    return v;
}
// two partial sums (each < 2^16 after reduction) per shuffle
__device__ __forceinline__ void red8x2( int &a, int &b, const Grp &g )
{
    u32 v = (u32)a | ( (u32)b << 16 );
    v += __shfl_xor_sync( g.mask, v, 1 );
    v += __shfl_xor_sync( g.mask, v, 2 );
    v += __shfl_xor_sync( g.mask, v, 4 );
    a = v & 0xffff; b = v >> 16;
}

// x264_pixel_satd_8x8 = satd_8x4(top) + satd_8x4(bottom) (pixel.c:290-330), rows across lanes
__device__ __forceinline__ int satd_rows( uint2 a, uint2 b, const Grp &g )
{
    u32 d[4], h0, h1, h2, h3;
    diff_row8( a, b, d );
    HADAMARD4( h0, h1, h2, h3, d[0], d[1], d[2], d[3] );
#define VSTAGE( v, bit ) { u32 o_ = __shfl_xor_sync( g.mask, v, bit ); v = ( g.r & bit ) ? o_ - v : v + o_; }
    VSTAGE( h0, 1 ) VSTAGE( h1, 1 ) VSTAGE( h2, 1 ) VSTAGE( h3, 1 )
    VSTAGE( h0, 2 ) VSTAGE( h1, 2 ) VSTAGE( h2, 2 ) VSTAGE( h3, 2 )
#undef VSTAGE
    u32 s = abs2( h0 ) + abs2( h1 ) + abs2( h2 ) + abs2( h3 );
    s += __shfl_xor_sync( g.mask, s, 1 );
    s += __shfl_xor_sync( g.mask, s, 2 );
    int v = (int)( ( ( s & 0xffff ) + ( s >> 16 ) ) >> 1 );
    v += __shfl_xor_sync( g.mask, v, 4 );
    return v;
}

__device__ __forceinline__ int cmp_rows( uint2 a, uint2 b, const Grp &g, int b_satd )
{
    return b_satd ? satd_rows( a, b, g ) : red8( (int)sad8( a, b ), g );
}

// get_ref (mc.c:221-249) for this lane's row: `ref` points at this lane's row of the F plane at mv 0
__device__ __forceinline__ uint2 qpel_row( const uint8_t *ref, const LaParams &P, int mvx, int mvy, int wscale, int wdenom, int woffset )
{
    int qidx = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
    int64_t off = (int64_t)( mvy >> 2 ) * P.stride + ( mvx >> 2 );
    const uint8_t *s1 = ref + c_hpel_ref0[qidx] * P.plane_size + off + ( ( mvy & 3 ) == 3 ? P.stride : 0 );
    uint2 v = load8u( s1 );
    if( qidx & 5 )
    {
        const uint8_t *s2 = ref + c_hpel_ref1[qidx] * P.plane_size + off + ( ( mvx & 3 ) == 3 );
        uint2 w = load8u( s2 );
        v.x = avg4( v.x, w.x ); v.y = avg4( v.y, w.y );
    }
    if( wdenom >= 0 )
    {
        v.x = weight4( v.x, wscale, wdenom, woffset );
        v.y = weight4( v.y, wscale, wdenom, woffset );
    }
    return v;
}

// pixel_avg_8x8 with weight (mc.c:63-99): (a*w + b*(64-w) + 32) >> 6 on 4 packed pixels.  The lookahead's
// weights are 64 - (dist_scale_factor>>2) in [1,63] (slicetype.c:529) or 32, a convex combination, so the clip of
// mc.c:86 never acts and two pixels fit one 16:16 multiply-add without carries between the halves.
__device__ __forceinline__ u32 biavg4( u32 a, u32 b, int w )
{
    if( w == 32 ) return avg4( a, b );
    // everything scaled by 4, so that each result byte lands on a byte boundary (bits 8..15 of its 16-bit field; the field's
    // maximum is 255 * 256 + 128 < 2^16): one PRMT then gathers the four bytes instead of two shifts, two masks and an OR
    const u32 w4 = 4u * (u32)w, v4 = 256u - w4;
    const u32 ae = __byte_perm( a, 0, 0x4240 ), ao = __byte_perm( a, 0, 0x4341 ), be = __byte_perm( b, 0, 0x4240 ), bo = __byte_perm( b, 0, 0x4341 );
    const u32 re = ae * w4 + be * v4 + 0x00800080u;      // pixels 0, 2
    const u32 ro = ao * w4 + bo * v4 + 0x00800080u;      // pixels 1, 3
    return __byte_perm( re, ro, 0x7351 );
}

// ------------------------------------------------------------------------------------------------
// motion search and mode decision
//
// The control flow (x264_me_search_ref, refine_subpel, the tail of slicetype_mb_cost) is written ONCE,
// templated on a cost "backend" that says how a block's pixels are spread over lanes:
//   GroupBE  -- 8 lanes per block (lane = pixel row): lowest latency, used when a launch has few
//               blocks (single evaluations through the per-call hook);
//   ThreadBE -- 1 thread per block, 32 horizontally adjacent blocks per warp: a warp-wide row load
//               touches 2-3 cache lines instead of 8 and nothing is recomputed per lane; used for
//               the wide launches of the batched path (ncu r01: the 8-lane form was L1TEX-wavefront
//               bound at 91 %).
// Both produce identical results (tests run each mapping against the oracle).
// ------------------------------------------------------------------------------------------------
struct MeGeo
{
    int wscale, wdenom, woffset;
    int mvpx, mvpy;
    int fminx, fminy, fmaxx, fmaxy;     // mv_limit_fpel
    int sminx, sminy, smaxx, smaxy;     // mv_min_spel / mv_max_spel
};

__device__ __forceinline__ void me_limits( MeGeo &m, const LaParams &P, int bx, int by )
{   // slicetype.c:550-562 (the y limits are refreshed at every row start, so they depend on by only)
    m.sminx = max( 4*( -8*bx - 12 ), -P.mv_range2 );
    m.smaxx = min( 4*( 8*( P.mbw - bx - 1 ) + 12 ), P.mv_range2 - 1 );
    m.sminy = max( 4*( -8*by - 12 ), -P.mv_range2 );
    m.smaxy = min( 4*( 8*( P.mbh - by - 1 ) + 12 ), P.mv_range2 - 1 );
    m.fminx = m.sminx >> 2; m.fmaxx = m.smaxx >> 2; m.fminy = m.sminy >> 2; m.fmaxy = m.smaxy >> 2;
}

// cost_mv table (analyse.c:153-157), read from global memory through L1: two 16-bit loads per candidate that issue
// ahead of the pixel loads.  (r01 history: a full per-CTA shared-memory copy cost 14 % of the search kernel's stall
// samples in the fill; a 512-entry copy with a global fallback cost ~25 instructions per candidate in address
// selection.)
struct CostTab
{
    const uint16_t *g;
    __device__ __forceinline__ int operator[]( int d ) const { return (int)__ldg( g + d ); }
};
__device__ __forceinline__ CostTab stage_cost_tab( const LaParams &P )
{
    CostTab t; t.g = P.cost_mv;
    return t;
}

__device__ __forceinline__ int mvcost( const CostTab &tab, const MeGeo &m, int mx, int my )
{   // p_cost_mvx[mx] + p_cost_mvy[my] (me.c:211-212); the table is symmetric (analyse.c:153-157)
    return tab[abs( mx - m.mvpx )] + tab[abs( my - m.mvpy )];
}
__device__ __forceinline__ bool in_range( const MeGeo &m, int mx, int my )
{   // CHECK_MVRANGE (me.c:209)
    return mx >= m.fminx && mx <= m.fmaxx && my >= m.fminy && my <= m.fmaxy;
}

// Row loader of ThreadBE: 8 pixels at byte offset `off` (any alignment) from an 8-byte aligned origin.
// LD32: three aligned words + two funnel shifts.  LD64: two aligned 64-bit loads cover the 8 pixels at any alignment
// (4 L1 wavefronts per warp instead of 6) at the price of three selects.  B200 measurements (profiles/r01): the search
// kernels gain from LD64, the decision kernels -- bound by L2->L1 fill traffic, which the wider window increases --
// lose 28 %, so it is a per-kernel choice.  The pointer is aligned ONCE per candidate; rows are immediates off it.
// LDM (load mode): 0 = three aligned 32-bit words; 1 = two aligned 64-bit words (16-byte window); 2 = one aligned 64-bit word + one
// 32-bit word: the same 12-byte window as mode 0 in two L1 requests instead of three -- the 64-bit load takes words (0,1) when
// the first word index is even and words (1,2) when it is odd, the 32-bit load takes the remaining one, three selects put
// them back in order.
template<int LDM>
struct RowSrc
{
    const u32 *q;       // mode 2: the 64-bit pointer
    const u32 *q3;      // mode 2 only: the 32-bit pointer
    u32 sh;
    int k;
    __device__ __forceinline__ RowSrc( const uint8_t *origin, int off )
    {
        if( LDM == 1 )      { q = (const u32 *)origin + ( ( off >> 3 ) << 1 ); k = off & 4; }
        else if( LDM == 2 ) { const int w = off >> 2; k = w & 1; q = (const u32 *)origin + w + k; q3 = (const u32 *)origin + w + ( k ? 0 : 2 ); }
        else                { q = (const u32 *)origin + ( off >> 2 ); k = 0; }
        sh = ( off & 3 ) * 8;
        // pin the row-0 pointer, the shift and the word select in registers: under register pressure the compiler
        // otherwise rematerialises the whole offset computation (plane table lookup, multiply, 64-bit add) for every row
        asm volatile( "" : "+l"( q ), "+r"( sh ), "+r"( k ) );
        if( LDM == 2 ) asm volatile( "" : "+l"( q3 ) );
    }
    __device__ __forceinline__ uint2 row( int w4 ) const      // w4: row offset in words (a multiple of 2)
    {
        if( LDM == 1 )
        {
            const uint2 A = __ldg( (const uint2 *)( q + w4 ) ), B = __ldg( (const uint2 *)( q + w4 ) + 1 );
            const u32 a = k ? A.y : A.x, b = k ? B.x : A.y, c = k ? B.y : B.x;
            return make_uint2( __funnelshift_r( a, b, sh ), __funnelshift_r( b, c, sh ) );
        }
        if( LDM == 2 )
        {
            const uint2 A = __ldg( (const uint2 *)( q + w4 ) ); const u32 c = __ldg( q3 + w4 );
            const u32 w0 = k ? c : A.x, w1 = k ? A.x : A.y, w2 = k ? A.y : c;
            return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
        }
        const u32 w0 = __ldg( q + w4 ), w1 = __ldg( q + w4 + 1 ), w2 = __ldg( q + w4 + 2 );
        return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
    }
    // 9 pixels of a row as three words (bytes 0..11 of the window; only byte 8 of the third word is guaranteed)
    __device__ __forceinline__ void row12( int w4, u32 &t0, u32 &t1, u32 &t2 ) const
    {
        if( LDM == 1 )
        {
            const uint2 A = __ldg( (const uint2 *)( q + w4 ) ), B = __ldg( (const uint2 *)( q + w4 ) + 1 );
            const u32 a = k ? A.y : A.x, b = k ? B.x : A.y, c = k ? B.y : B.x, d = k ? B.y : B.y;
            t0 = __funnelshift_r( a, b, sh ); t1 = __funnelshift_r( b, c, sh ); t2 = __funnelshift_r( c, d, sh );
            return;
        }
        if( LDM == 2 )
        {
            const uint2 A = __ldg( (const uint2 *)( q + w4 ) ); const u32 c = __ldg( q3 + w4 );
            const u32 w0 = k ? c : A.x, w1 = k ? A.x : A.y, w2 = k ? A.y : c;
            t0 = __funnelshift_r( w0, w1, sh ); t1 = __funnelshift_r( w1, w2, sh ); t2 = w2 >> sh;
            return;
        }
        const u32 w0 = __ldg( q + w4 ), w1 = __ldg( q + w4 + 1 ), w2 = __ldg( q + w4 + 2 ), w3 = __ldg( q + w4 + 3 );
        t0 = __funnelshift_r( w0, w1, sh ); t1 = __funnelshift_r( w1, w2, sh ); t2 = __funnelshift_r( w2, w3, sh );
    }
    // The same fetch under a per-thread predicate, written in PTX so that it stays predicated: get_ref reads a second
    // plane only at quarter-pel positions, and whether the compiler turns `if( avg )` into predication or into two
    // specialised copies of the 8-row body flips with unrelated edits (r01: the branchy form runs 13 % slower, because
    // a warp whose blocks disagree on `avg` executes both copies).  When p is 0 nothing is loaded and junk derived
    // from `v` comes back (the caller discards it).
    __device__ __forceinline__ uint2 row_if( int p, int w4, uint2 v ) const
    {
        if( LDM == 1 )
        {
            u32 a0 = v.x, a1 = v.y, b0 = v.x, b1 = v.y;
            asm( "{ .reg .pred p;\n setp.ne.s32 p, %4, 0;\n @p ld.global.nc.v2.u32 {%0,%1}, [%5];\n @p ld.global.nc.v2.u32 {%2,%3}, [%5+8];\n}"
                 : "+r"( a0 ), "+r"( a1 ), "+r"( b0 ), "+r"( b1 ) : "r"( p ), "l"( q + w4 ) );
            const u32 a = k ? a1 : a0, b = k ? b0 : a1, c = k ? b1 : b0;
            return make_uint2( __funnelshift_r( a, b, sh ), __funnelshift_r( b, c, sh ) );
        }
        if( LDM == 2 )
        {
#if X264CU_ROWIF_ZERO
            u32 a0 = 0, a1 = 0, c = 0;          // nothing that depends on the first plane's data: the loads can leave at once
#else
            u32 a0 = v.x, a1 = v.y, c = v.x;
#endif
            asm( "{ .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p ld.global.nc.v2.u32 {%0,%1}, [%4];\n @p ld.global.nc.u32 %2, [%5];\n}"
                 : "+r"( a0 ), "+r"( a1 ), "+r"( c ) : "r"( p ), "l"( q + w4 ), "l"( q3 + w4 ) );
            const u32 w0 = k ? c : a0, w1 = k ? a0 : a1, w2 = k ? a1 : c;
            return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
        }
        u32 w0 = v.x, w1 = v.y, w2 = v.x;
        asm( "{ .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p ld.global.nc.u32 %0, [%4];\n @p ld.global.nc.u32 %1, [%4+4];\n @p ld.global.nc.u32 %2, [%4+8];\n}"
             : "+r"( w0 ), "+r"( w1 ), "+r"( w2 ) : "r"( p ), "l"( q + w4 ) );
        return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
    }
    // row12 under a per-thread predicate (see row_if)
    __device__ __forceinline__ void row12_if( int p, int w4, u32 &t0, u32 &t1, u32 &t2 ) const
    {
        static_assert( LDM == 2, "row12_if: mode 2 only" );
        u32 a0 = 0, a1 = 0, c = 0;
        asm( "{ .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p ld.global.nc.v2.u32 {%0,%1}, [%4];\n @p ld.global.nc.u32 %2, [%5];\n}"
             : "+r"( a0 ), "+r"( a1 ), "+r"( c ) : "r"( p ), "l"( q + w4 ), "l"( q3 + w4 ) );
        const u32 w0 = k ? c : a0, w1 = k ? a0 : a1, w2 = k ? a1 : c;
        t0 = __funnelshift_r( w0, w1, sh ); t1 = __funnelshift_r( w1, w2, sh ); t2 = w2 >> sh;
    }
};

static __constant__ int8_t c_hex2[8][2] = {{-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0}};   // me.c:55
// the same table as two immediates (4 bits per entry, value + 2): no indexed constant load on the candidate path
__device__ __forceinline__ int hex2x( int i ) { return (int)( ( 0x01343101u >> ( 4*i ) ) & 7u ) - 2; }
__device__ __forceinline__ int hex2y( int i ) { return (int)( ( 0x20024420u >> ( 4*i ) ) & 7u ) - 2; }
static __constant__ int8_t c_sq1[9][2] = {{0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1}}; // me.c:56
__device__ __forceinline__ int sq1x( int i ) { return (int)( ( 0x28215u >> ( 2*i ) ) & 3u ) - 1; }      // the same as immediates
__device__ __forceinline__ int sq1y( int i ) { return (int)( ( 0x22161u >> ( 2*i ) ) & 3u ) - 1; }

// ---- 8 lanes per block -------------------------------------------------------------------------------
struct GroupBE
{
    static constexpr bool PAR4 = false, HPEL_EVEN = false, SATD5 = false;
    Grp g;
    uint2 fenc;              // this lane's row of the block
    const uint8_t *ref0;     // this lane's row of reference 0's F plane at mv 0
    const uint8_t *ref1;     // same for reference 1 (decision kernel only)
    const uint8_t *ref_w;    // this lane's row of the fullpel (possibly weighted) plane of reference 0
    const LaParams *P;

    __device__ __forceinline__ bool leader() const { return g.r == 0; }
    __device__ __forceinline__ bool satd() const { return P->b_satd != 0; }
    __device__ __forceinline__ int zero_cmp() const { return cmp_rows( fenc, load8u( ref0 ), g, P->b_satd ); }
    // candidates k = 0..n-1 in order, strict '<' (the reference's COPY*_IF_LT tie-breaking); two candidates share
    // one 3-shuffle reduction (16:16 packed partial sums)
    template<class XY> __device__ __forceinline__ void fpel_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        for( int k = 0; k < n; k += 2 )
        {
            int x0, y0, x1, y1;
            xy( k, x0, y0 ); xy( min( k+1, n-1 ), x1, y1 );
            int c0 = (int)sad8( fenc, load8u( ref_w + (int64_t)y0 * P->stride + x0 ) );
            int c1 = (int)sad8( fenc, load8u( ref_w + (int64_t)y1 * P->stride + x1 ) );
            red8x2( c0, c1, g );
            c0 += mvcost( tab, m, x0*4, y0*4 ); c1 += mvcost( tab, m, x1*4, y1*4 );
            if( c0 < bcost ) { bcost = c0; best = k; }
            if( k+1 < n && c1 < bcost ) { bcost = c1; best = k+1; }
        }
    }
    template<class XY> __device__ __forceinline__ void qsad_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        for( int k = 0; k < n; k += 2 )
        {
            int x0, y0, x1, y1;
            xy( k, x0, y0 ); xy( min( k+1, n-1 ), x1, y1 );
            int c0 = (int)sad8( fenc, qpel_row( ref0, *P, x0, y0, m.wscale, m.wdenom, m.woffset ) );
            int c1 = (int)sad8( fenc, qpel_row( ref0, *P, x1, y1, m.wscale, m.wdenom, m.woffset ) );
            red8x2( c0, c1, g );
            c0 += mvcost( tab, m, x0, y0 ); c1 += mvcost( tab, m, x1, y1 );
            if( c0 < bcost ) { bcost = c0; best = k; }
            if( k+1 < n && c1 < bcost ) { bcost = c1; best = k+1; }
        }
    }
    __device__ __forceinline__ void hex6( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    { fpel_best( 6, [&]( int k, int &x, int &y ) { x = ox + c_hex2[k+1][0]; y = oy + c_hex2[k+1][1]; }, tab, m, bcost, best ); }
    __device__ __forceinline__ void square8( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    { fpel_best( 8, [&]( int k, int &x, int &y ) { x = ox + c_sq1[k+1][0]; y = oy + c_sq1[k+1][1]; }, tab, m, bcost, best ); }
    __device__ __forceinline__ int fpel1( int mx, int my ) const { return red8( (int)sad8( fenc, load8u( ref_w + (int64_t)my * P->stride + mx ) ), g ); }
    __device__ __forceinline__ int qsad1( const MeGeo &m, int mx, int my ) const
    { return red8( (int)sad8( fenc, qpel_row( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) ), g ); }
    __device__ __forceinline__ int qsatd( const MeGeo &m, int mx, int my ) const
    {
        return satd_rows( fenc, qpel_row( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ), g );
    }
    // TRY_BIDIR without the penalty (slicetype.c:579-603)
    __device__ __forceinline__ int bidir( int mv0x, int mv0y, int mv1x, int mv1y, int w ) const
    {
        uint2 a, b;
        if( !P->b_satd )
        {   // subme <= 1: hpel planes addressed directly (slicetype.c:582-589)
            int i0 = ( ( mv0x & 2 ) >> 1 ) + ( mv0y & 2 ), i1 = ( ( mv1x & 2 ) >> 1 ) + ( mv1y & 2 );
            a = load8u( ref0 + i0*P->plane_size + ( mv0x >> 2 ) + (int64_t)( mv0y >> 2 ) * P->stride );
            b = load8u( ref1 + i1*P->plane_size + ( mv1x >> 2 ) + (int64_t)( mv1y >> 2 ) * P->stride );
        }
        else
        {
            a = qpel_row( ref0, *P, mv0x, mv0y, 0, -1, 0 );
            b = qpel_row( ref1, *P, mv1x, mv1y, 0, -1, 0 );
        }
        return cmp_rows( fenc, make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ), g, P->b_satd );
    }
    __device__ __forceinline__ int bidir_zero( int w ) const
    {   // slicetype.c:648-649
        uint2 a = load8u( ref0 ), b = load8u( ref1 );
        return cmp_rows( fenc, make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ), g, P->b_satd );
    }
};

#ifndef X264CU_SPIN_NS0
#define X264CU_SPIN_NS0 20
#define X264CU_SPIN_NS1 640
#endif
// ---- 32 lanes per block: 4 candidates x 8 pixel rows ---------------------------------------------------
// The latency-oriented mapping of the exact raster-serial mode (search_kernel_serial): every candidate round of
// x264_me_search scores up to FOUR candidates at once, one per 8-lane subgroup (lane & 7 = pixel row, lane >> 3 =
// candidate slot); the winner is the minimum of (cost << 2 | slot) over the subgroups, which is the reference's
// "strict <, earlier candidate wins ties" rule.  Single evaluations are computed redundantly by the four subgroups.
template<bool SATD>
struct WarpBE
{
    static constexpr bool PAR4 = true, HPEL_EVEN = false, SATD5 = false;
    Grp g;                   // mask: the whole warp; r: pixel row
    int sub;                 // candidate slot of this lane
    uint2 fenc;
    const uint8_t *ref0, *ref1, *ref_w;      // this lane's row at mv 0
    const LaParams *P;
    int *err;                // error flag of the launch (bounded spins)

    // The three MV words of the row below (below, below-left, below-right; `need` bit i: word i exists): spin with L1-bypassing
    // loads until their producers have replaced the "not ready" pattern.  Three lanes poll at once: one L2 round trip on the serial critical path instead of three (a word that its producer has already stored --
    // the common case, the row below runs two blocks ahead -- comes back on the first try).
    __device__ __forceinline__ void wait_mv3( const int *below, int need, int &v0, int &v1, int &v2 ) const
    {
        const int lane = threadIdx.x & 31;
        const int *p = below + ( lane == 1 ? -1 : lane == 2 ? 1 : 0 );
        const bool mine = lane < 3 && ( ( need >> lane ) & 1 );
        int v = 0;
        unsigned ns = X264CU_SPIN_NS0;
        for( int spins = 1; ; spins++ )
        {
            if( mine ) asm volatile( "ld.volatile.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
            const bool pending = mine && v == 0x7F7F7F7F;
            if( !__any_sync( 0xffffffffu, pending ) ) break;
            if( ( spins & 1023 ) == 0 && ( *(volatile int *)err || spins > ( 1 << 22 ) ) ) { *(volatile int *)err = 1; if( pending ) v = 0; break; }
            __nanosleep( ns );
            ns = min( ns * 2, (unsigned)X264CU_SPIN_NS1 );
        }
        v0 = __shfl_sync( 0xffffffffu, v, 0 ); v1 = __shfl_sync( 0xffffffffu, v, 1 ); v2 = __shfl_sync( 0xffffffffu, v, 2 );
    }
    // sum over the 8 lanes of this lane's subgroup: xor-shuffles (REDUX with four different 8-lane masks per warp
    // measured 60 % slower, profiles/r01/tuning.md)
    __device__ __forceinline__ int rsum8( int v ) const { return red8( v, g ); }
    // x264_pixel_satd_8x8 with the rows across the subgroup's lanes
    __device__ __forceinline__ int satd8( uint2 a, uint2 b ) const
    {
        u32 d[4], h0, h1, h2, h3;
        diff_row8( a, b, d );
        HADAMARD4( h0, h1, h2, h3, d[0], d[1], d[2], d[3] );
#define VSTAGE( v, bit ) { u32 o_ = __shfl_xor_sync( 0xffffffffu, v, bit ); v = ( g.r & bit ) ? o_ - v : v + o_; }
        VSTAGE( h0, 1 ) VSTAGE( h1, 1 ) VSTAGE( h2, 1 ) VSTAGE( h3, 1 )
        VSTAGE( h0, 2 ) VSTAGE( h1, 2 ) VSTAGE( h2, 2 ) VSTAGE( h3, 2 )
#undef VSTAGE
        const u32 s = abs2( h0 ) + abs2( h1 ) + abs2( h2 ) + abs2( h3 );
        // each 8x4 half: ((sum & 0xffff) + (sum >> 16)) >> 1 over its 4 lanes (pixel.c:300-306)
        u32 t = s + __shfl_xor_sync( 0xffffffffu, s, 1 );
        t += __shfl_xor_sync( 0xffffffffu, t, 2 );
        const int v = (int)( ( ( t & 0xffff ) + ( t >> 16 ) ) >> 1 );
        return v + __shfl_xor_sync( 0xffffffffu, v, 4 );
    }
    __device__ __forceinline__ int cmp8( uint2 a, uint2 b ) const { return SATD ? satd8( a, b ) : rsum8( (int)sad8( a, b ) ); }
    __device__ __forceinline__ bool leader() const { return ( threadIdx.x & 31 ) == 0; }
    __device__ __forceinline__ bool satd() const { return SATD; }
    __device__ __forceinline__ int zero_cmp() const { return cmp8( fenc, load8u( ref0 ) ); }
    // slot s scores candidate k0 + s; `c` is this subgroup's full cost (identical in its 8 lanes)
    __device__ __forceinline__ void take_min4( int c, bool valid, int k0, int &bcost, int &best ) const
    {
        u32 key = valid ? ( (u32)c << 2 ) | (u32)sub : 0xffffffffu;
        key = __reduce_min_sync( 0xffffffffu, key );                 // REDUX over the whole warp: one instruction, 7 % faster than two shuffle stages
        if( key != 0xffffffffu && (int)( key >> 2 ) < bcost ) { bcost = (int)( key >> 2 ); best = k0 + (int)( key & 3 ); }
    }
    template<class XY> __device__ __forceinline__ void fpel_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        for( int k0 = 0; k0 < n; k0 += 4 )
        {
            int x, y;
            xy( min( k0 + sub, n-1 ), x, y );
            const int c = rsum8( (int)sad8( fenc, load8u( ref_w + (int64_t)y * P->stride + x ) ) ) + mvcost( tab, m, x*4, y*4 );
            take_min4( c, k0 + sub < n, k0, bcost, best );
        }
    }
    template<class XY> __device__ __forceinline__ void qsad_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        for( int k0 = 0; k0 < n; k0 += 4 )
        {
            int x, y;
            xy( min( k0 + sub, n-1 ), x, y );
            const int c = rsum8( (int)sad8( fenc, qpel_row( ref0, *P, x, y, m.wscale, m.wdenom, m.woffset ) ) ) + mvcost( tab, m, x, y );
            take_min4( c, k0 + sub < n, k0, bcost, best );
        }
    }
    // qsad_best that also reports the cost of candidate 0 (the predictor of me.c:216-275)
    template<class XY> __device__ __forceinline__ void qsad_best0( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best, int &cost0 ) const
    {
        for( int k0 = 0; k0 < n; k0 += 4 )
        {
            int x, y;
            xy( min( k0 + sub, n-1 ), x, y );
            const int c = rsum8( (int)sad8( fenc, qpel_row( ref0, *P, x, y, m.wscale, m.wdenom, m.woffset ) ) ) + mvcost( tab, m, x, y );
            if( k0 == 0 ) cost0 = __shfl_sync( 0xffffffffu, c, 0 );
            take_min4( c, k0 + sub < n, k0, bcost, best );
        }
    }
    // four SATD candidates at once; bit k of `skip` drops candidate k (me.c:947-963)
    template<class XY> __device__ __forceinline__ void qsatd_best4( XY xy, int skip, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        int x, y;
        xy( sub, x, y );
        const int c = satd8( fenc, qpel_row( ref0, *P, x, y, m.wscale, m.wdenom, m.woffset ) ) + mvcost( tab, m, x, y );
        take_min4( c, !( ( skip >> sub ) & 1 ), 0, bcost, best );
    }
    __device__ __forceinline__ void hex6( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    { fpel_best( 6, [&]( int k, int &x, int &y ) { x = ox + hex2x( k+1 ); y = oy + hex2y( k+1 ); }, tab, m, bcost, best ); }
    __device__ __forceinline__ void square8( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    { fpel_best( 8, [&]( int k, int &x, int &y ) { x = ox + c_sq1[k+1][0]; y = oy + c_sq1[k+1][1]; }, tab, m, bcost, best ); }
    __device__ __forceinline__ int fpel1( int mx, int my ) const { return rsum8( (int)sad8( fenc, load8u( ref_w + (int64_t)my * P->stride + mx ) ) ); }
    __device__ __forceinline__ int qsad1( const MeGeo &m, int mx, int my ) const
    { return rsum8( (int)sad8( fenc, qpel_row( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) ) ); }
    __device__ __forceinline__ int qsatd( const MeGeo &m, int mx, int my ) const
    { return satd8( fenc, qpel_row( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) ); }
};

// ---- 32 lanes per block: 8 candidates x 4 lanes x 2 pixel rows ------------------------------------------
// The second latency-oriented mapping of the exact raster-serial mode (tuning build, -DX264CU_SERIAL_SLOTS=8).  A block search is a chain of candidate rounds, each
// a dependent sequence of address arithmetic, loads, the metric, a shuffle reduction and the warp-wide minimum (~500 cycles);
// with four slots the full hexagon, the square refine and the predictor set take two rounds each and the SATD re-measurement
// of the winner is a round of its own.  Eight slots (lane >> 2 = slot, lane & 3 = row pair) make each of them ONE round --
// the re-measurement shares its round with the four quarter-pel neighbours -- and the per-slot reductions are two shuffle
// levels instead of three; every lane carries two rows, which adds instructions but not chain length.
template<bool SATD>
struct WarpBE8
{
    static constexpr bool PAR4 = true, HPEL_EVEN = false, SATD5 = true;
    static constexpr int SLOTS = 8;
    int sub, q;              // candidate slot of this lane; row pair: this lane holds rows 2q and 2q+1
    uint2 fenc0, fenc1;
    const uint8_t *ref0, *ref1, *ref_w;      // this lane's FIRST row at mv 0
    const LaParams *P;
    int *err;

    __device__ __forceinline__ void wait_mv3( const int *below, int need, int &v0, int &v1, int &v2 ) const
    {
        const int lane = threadIdx.x & 31;
        const int *p = below + ( lane == 1 ? -1 : lane == 2 ? 1 : 0 );
        const bool mine = lane < 3 && ( ( need >> lane ) & 1 );
        int v = 0;
        unsigned ns = X264CU_SPIN_NS0;
        for( int spins = 1; ; spins++ )
        {
            if( mine ) asm volatile( "ld.volatile.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
            const bool pending = mine && v == 0x7F7F7F7F;
            if( !__any_sync( 0xffffffffu, pending ) ) break;
            if( ( spins & 1023 ) == 0 && ( *(volatile int *)err || spins > ( 1 << 22 ) ) ) { *(volatile int *)err = 1; if( pending ) v = 0; break; }
            __nanosleep( ns );
            ns = min( ns * 2, (unsigned)X264CU_SPIN_NS1 );
        }
        v0 = __shfl_sync( 0xffffffffu, v, 0 ); v1 = __shfl_sync( 0xffffffffu, v, 1 ); v2 = __shfl_sync( 0xffffffffu, v, 2 );
    }
    __device__ __forceinline__ bool leader() const { return ( threadIdx.x & 31 ) == 0; }
    __device__ __forceinline__ bool satd() const { return SATD; }
    // sum over the 4 lanes of this lane's slot
    __device__ __forceinline__ int rsum4( int v ) const
    {
        v += __shfl_xor_sync( 0xffffffffu, v, 1 );
        return v + __shfl_xor_sync( 0xffffffffu, v, 2 );
    }
    __device__ __forceinline__ int sad2( uint2 a0, uint2 a1 ) const { return rsum4( (int)( sad8( fenc0, a0 ) + sad8( fenc1, a1 ) ) ); }
    // x264_pixel_satd_8x8 (pixel.c:290-330): lanes q = 0,1 hold the upper 8x4 half, q = 2,3 the lower one.  Per column word the
    // first vertical butterfly (row 2q +- row 2q+1) is in registers; the second one, between the two lanes of a half, is folded
    // into the absolute sum: |u+v| + |u-v| == 2*max(|u|,|v|) per 16-bit field, so a lane only needs its partner's absolute
    // values -- and only for the two column words it scores (even q: words 0,1; odd q: words 2,3).
    __device__ __forceinline__ int satd2( uint2 a0, uint2 a1 ) const
    {
        u32 d[4], h0[4], h1[4];
        diff_row8( fenc0, a0, d ); HADAMARD4( h0[0], h0[1], h0[2], h0[3], d[0], d[1], d[2], d[3] );
        diff_row8( fenc1, a1, d ); HADAMARD4( h1[0], h1[1], h1[2], h1[3], d[0], d[1], d[2], d[3] );
        u32 as[4], ad[4];
#pragma unroll
        for( int i = 0; i < 4; i++ ) { as[i] = abs2( h0[i] + h1[i] ); ad[i] = abs2( h0[i] - h1[i] ); }
        const bool odd = q & 1;
        // send the two words the partner scores, receive the partner's values for the two words this lane scores
        const u32 ps0 = __shfl_xor_sync( 0xffffffffu, odd ? as[0] : as[2], 1 ), ps1 = __shfl_xor_sync( 0xffffffffu, odd ? as[1] : as[3], 1 );
        const u32 pd0 = __shfl_xor_sync( 0xffffffffu, odd ? ad[0] : ad[2], 1 ), pd1 = __shfl_xor_sync( 0xffffffffu, odd ? ad[1] : ad[3], 1 );
        const u32 ms0 = odd ? as[2] : as[0], ms1 = odd ? as[3] : as[1], md0 = odd ? ad[2] : ad[0], md1 = odd ? ad[3] : ad[1];
        const u32 sum = __vmaxu2( ms0, ps0 ) + __vmaxu2( ms1, ps1 ) + __vmaxu2( md0, pd0 ) + __vmaxu2( md1, pd1 );
        return rsum4( (int)( ( sum & 0xffff ) + ( sum >> 16 ) ) );
    }
    __device__ __forceinline__ int cmp2( uint2 a0, uint2 a1 ) const { return SATD ? satd2( a0, a1 ) : sad2( a0, a1 ); }
    __device__ __forceinline__ int zero_cmp() const { return cmp2( load8u( ref0 ), load8u( ref0 + P->stride ) ); }
    __device__ __forceinline__ int fpel_cost( int x, int y ) const
    {
        const uint8_t *p = ref_w + (int64_t)y * P->stride + x;
        return sad2( load8u( p ), load8u( p + P->stride ) );
    }
    __device__ __forceinline__ int qsad_cost( const MeGeo &m, int x, int y ) const
    {
        return sad2( qpel_row( ref0, *P, x, y, m.wscale, m.wdenom, m.woffset ), qpel_row( ref0 + P->stride, *P, x, y, m.wscale, m.wdenom, m.woffset ) );
    }
    __device__ __forceinline__ int qsatd_cost( const MeGeo &m, int x, int y ) const
    {
        return satd2( qpel_row( ref0, *P, x, y, m.wscale, m.wdenom, m.woffset ), qpel_row( ref0 + P->stride, *P, x, y, m.wscale, m.wdenom, m.woffset ) );
    }
    // slot s scores candidate k0 + s; `c` is this slot's full cost (identical in its 4 lanes).  The minimum of (cost << 3 | slot)
    // is the reference's "strict <, earlier candidate wins ties" rule.
    __device__ __forceinline__ void take_min8( int c, bool valid, int k0, int &bcost, int &best ) const
    {
        u32 key = valid ? ( (u32)c << 3 ) | (u32)sub : 0xffffffffu;
        key = __reduce_min_sync( 0xffffffffu, key );
        if( key != 0xffffffffu && (int)( key >> 3 ) < bcost ) { bcost = (int)( key >> 3 ); best = k0 + (int)( key & 7 ); }
    }
    template<class XY> __device__ __forceinline__ void fpel_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        for( int k0 = 0; k0 < n; k0 += SLOTS )
        {
            int x, y;
            xy( min( k0 + sub, n-1 ), x, y );
            const int c = fpel_cost( x, y ) + mvcost( tab, m, x*4, y*4 );
            take_min8( c, k0 + sub < n, k0, bcost, best );
        }
    }
    template<class XY> __device__ __forceinline__ void qsad_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        for( int k0 = 0; k0 < n; k0 += SLOTS )
        {
            int x, y;
            xy( min( k0 + sub, n-1 ), x, y );
            const int c = qsad_cost( m, x, y ) + mvcost( tab, m, x, y );
            take_min8( c, k0 + sub < n, k0, bcost, best );
        }
    }
    template<class XY> __device__ __forceinline__ void qsad_best0( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best, int &cost0 ) const
    {
        for( int k0 = 0; k0 < n; k0 += SLOTS )
        {
            int x, y;
            xy( min( k0 + sub, n-1 ), x, y );
            const int c = qsad_cost( m, x, y ) + mvcost( tab, m, x, y );
            if( k0 == 0 ) cost0 = __shfl_sync( 0xffffffffu, c, 0 );
            take_min8( c, k0 + sub < n, k0, bcost, best );
        }
    }
    // The SATD re-measurement of the winner (me.c:925-929, slot 0) and, when `diamond`, its four quarter-pel neighbours
    // (me.c:947-963, slots 1..4) in one round: bcost becomes the centre's SATD cost unless a neighbour is strictly cheaper;
    // best = -1: the centre stays, else the neighbour index 0..3 (up, down, left, right).
    __device__ __forceinline__ void qsatd_best5( int omx, int omy, bool diamond, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        const int k = min( sub, 4 ) - 1;
        const int x = omx + ( k == 2 ? -1 : k == 3 ? 1 : 0 ), y = omy + ( k == 0 ? -1 : k == 1 ? 1 : 0 );
        const int c = qsatd_cost( m, x, y ) + mvcost( tab, m, x, y );
        u32 key = ( sub == 0 || ( diamond && sub <= 4 ) ) ? ( (u32)c << 3 ) | (u32)sub : 0xffffffffu;
        key = __reduce_min_sync( 0xffffffffu, key );
        bcost = (int)( key >> 3 ); best = (int)( key & 7 ) - 1;
    }
    __device__ __forceinline__ void hex6( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    { fpel_best( 6, [&]( int k, int &x, int &y ) { x = ox + hex2x( k+1 ); y = oy + hex2y( k+1 ); }, tab, m, bcost, best ); }
    __device__ __forceinline__ void square8( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    { fpel_best( 8, [&]( int k, int &x, int &y ) { x = ox + sq1x( k+1 ); y = oy + sq1y( k+1 ); }, tab, m, bcost, best ); }
    __device__ __forceinline__ int fpel1( int mx, int my ) const { return fpel_cost( mx, my ); }
    __device__ __forceinline__ int qsad1( const MeGeo &m, int mx, int my ) const { return qsad_cost( m, mx, my ); }
    __device__ __forceinline__ int qsatd( const MeGeo &m, int mx, int my ) const { return qsatd_cost( m, mx, my ); }
};

// ---- 1 thread per block ---------------------------------------------------------------------------
// PITCH is the device plane pitch as a compile-time constant (0: read it from LaParams).  With a constant pitch the
// eight row loads of a candidate are one base address plus immediates; with a run-time pitch every row costs two
// 64-bit address instructions (ncu r01: address arithmetic was ~40 % of the executed instructions).
template<class RowFn>
__device__ __forceinline__ int sad8x8_t( const uint2 ( &fenc )[8], RowFn pred )
{
    u32 s0 = 0, s1 = 0;
#pragma unroll
    for( int r = 0; r < 8; r++ ) { uint2 p = pred( r ); s0 = __vsadu4( fenc[r].x, p.x ) + s0; s1 = __vsadu4( fenc[r].y, p.y ) + s1; }
    return (int)( s0 + s1 );
}
// x264_pixel_satd_8x8 = two satd_8x4 (pixel.c:290-330), all in registers
template<class RowFn>
__device__ __forceinline__ int satd8x8_t( const uint2 ( &fenc )[8], RowFn pred )
{
    int total = 0;
#pragma unroll
    for( int h = 0; h < 2; h++ )
    {
        u32 t[4][4];
#pragma unroll
        for( int i = 0; i < 4; i++ )
        {
            u32 d[4];
            diff_row8( fenc[4*h+i], pred( 4*h+i ), d );
            HADAMARD4( t[i][0], t[i][1], t[i][2], t[i][3], d[0], d[1], d[2], d[3] );
        }
        // second half of the vertical transform folded into the absolute sum: |u+v| + |u-v| == 2*max(|u|,|v|)
        // per 16-bit half (lane sums stay below 2^15 for 8-bit pixels), so the result equals pixel.c:300-306's
        u32 sum = 0;
#pragma unroll
        for( int i = 0; i < 4; i++ )
        {
            const u32 u0 = t[0][i] + t[1][i], u1 = t[0][i] - t[1][i], u2 = t[2][i] + t[3][i], u3 = t[2][i] - t[3][i];
            sum += __vmaxu2( abs2( u0 ), abs2( u2 ) ) + __vmaxu2( abs2( u1 ), abs2( u3 ) );
        }
        total += (int)( ( sum & 0xffff ) + ( sum >> 16 ) );
    }
    return total;
}

// 12 consecutive pixels starting at word pointer q (aligned) + sh bits, as three words
__device__ __forceinline__ void ld12w( const u32 *q, u32 sh, u32 &a0, u32 &a1, u32 &a2 )
{
    u32 w0 = __ldg( q ), w1 = __ldg( q + 1 ), w2 = __ldg( q + 2 ), w3 = __ldg( q + 3 );
    a0 = __funnelshift_r( w0, w1, sh ); a1 = __funnelshift_r( w1, w2, sh ); a2 = __funnelshift_r( w2, w3, sh );
}

// get_ref (mc.c:221-249) as a row generator: plane/offset selection hoisted out of the row loop
template<bool WEIGHTED, int PITCH, int LD64, bool PTXPRED>
struct QpelRows
{
    RowSrc<LD64> s1, s2;     // the two source planes at row 0
    int stride4;             // plane pitch in words
    int avg;
    int wscale, wdenom, woffset;
    __device__ __forceinline__ static int off1( const LaParams &P, int mvx, int mvy )
    {
        const int pitch = PITCH ? PITCH : P.stride;
        return ( mvy >> 2 ) * pitch + ( mvx >> 2 ) + ( ( mvy & 3 ) == 3 ? pitch : 0 ) + hpel_ref0( ( ( mvy & 3 ) << 2 ) + ( mvx & 3 ) ) * (int)P.plane_size;
    }
    __device__ __forceinline__ static int off2( const LaParams &P, int mvx, int mvy )
    {
        const int pitch = PITCH ? PITCH : P.stride;
        return ( mvy >> 2 ) * pitch + ( mvx >> 2 ) + ( ( mvx & 3 ) == 3 ) + hpel_ref1( ( ( mvy & 3 ) << 2 ) + ( mvx & 3 ) ) * (int)P.plane_size;
    }
    // ref: 8-byte aligned block origin in the F plane; plane_size and the pitch are multiples of 8
    __device__ __forceinline__ QpelRows( const uint8_t *ref, const LaParams &P, int mvx, int mvy, int ws, int wd, int wo )
        : s1( ref, off1( P, mvx, mvy ) ), s2( ref, off2( P, mvx, mvy ) )
    {
        stride4 = ( PITCH ? PITCH : P.stride ) >> 2;
        avg = ( ( ( ( mvy & 3 ) << 2 ) + ( mvx & 3 ) ) & 5 ) != 0;
        wscale = ws; wdenom = wd; woffset = wo;
    }
    __device__ __forceinline__ uint2 operator()( int r ) const
    {
        const int s4 = PITCH ? PITCH >> 2 : stride4;
        uint2 v = s1.row( r*s4 );
        if( PTXPRED )
        {
            const uint2 w = s2.row_if( avg, r*s4, v );
            v.x = avg ? avg4( v.x, w.x ) : v.x; v.y = avg ? avg4( v.y, w.y ) : v.y;
        }
#ifdef X264CU_QPEL_BOTH
        else if( true )
        {   // tuning: always fetch the second plane (its address is valid for every phase), select afterwards -- no branch, no predicate
            const uint2 w = s2.row( r*s4 );
            v.x = avg ? avg4( v.x, w.x ) : v.x; v.y = avg ? avg4( v.y, w.y ) : v.y;
        }
#endif
        else if( avg )
        {
            const uint2 w = s2.row( r*s4 );
            v.x = avg4( v.x, w.x ); v.y = avg4( v.y, w.y );
        }
        if( WEIGHTED && wdenom >= 0 ) { v.x = weight4( v.x, wscale, wdenom, woffset ); v.y = weight4( v.y, wscale, wdenom, woffset ); }
        return v;
    }
};

template<bool SATD, bool WEIGHTED, int PITCH, int LD64, bool PTXPRED>
struct ThreadBE
{
    static constexpr bool PAR4 = false, HPEL_EVEN = true, SATD5 = false;
    uint2 fenc[8];
    const uint8_t *ref0, *ref1, *ref_w;      // block origins (8-byte aligned)
    const LaParams *P;

    __device__ __forceinline__ int pitch() const { return PITCH ? PITCH : P->stride; }
    __device__ __forceinline__ bool leader() const { return true; }
    __device__ __forceinline__ bool satd() const { return SATD; }
    __device__ __forceinline__ int zero_cmp() const
    {
        const u32 *q = (const u32 *)ref0; const int s4 = pitch() >> 2;
        auto rows = [&]( int r ) { return __ldg( (const uint2 *)( q + r*s4 ) ); };
        return SATD ? satd8x8_t( fenc, rows ) : sad8x8_t( fenc, rows );
    }
    __device__ __forceinline__ int fpel1( int mx, int my ) const
    {
        const int s4 = pitch() >> 2;
        const RowSrc<LD64> src( ref_w, my * pitch() + mx );
        return sad8x8_t( fenc, [&]( int r ) { return src.row( r*s4 ); } );
    }
    __device__ __forceinline__ int qsad1( const MeGeo &m, int mx, int my ) const
    { return sad8x8_t( fenc, QpelRows<WEIGHTED, PITCH, LD64, PTXPRED>( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) ); }
    // candidate loops are NOT unrolled: one copy of the 8-row body per call site keeps the kernel inside the
    // instruction cache (ncu r01: the fully unrolled form had 25k SASS lines and 53 % no-instruction stalls)
    template<class XY> __device__ __forceinline__ void fpel_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
#pragma unroll 1
        for( int k = 0; k < n; k++ )
        {
            int x, y;
            xy( k, x, y );
            int c = fpel1( x, y ) + mvcost( tab, m, x*4, y*4 );
            if( c < bcost ) { bcost = c; best = k; }
        }
    }
    template<class XY> __device__ __forceinline__ void qsad_best( int n, XY xy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
#pragma unroll 1
        for( int k = 0; k < n; k++ )
        {
            int x, y;
            xy( k, x, y );
            int c = qsad1( m, x, y ) + mvcost( tab, m, x, y );
            if( c < bcost ) { bcost = c; best = k; }
        }
    }
    // The half-pel diamond (me.c:898-921) around a centre whose components are both even (full- or half-pel): the four
    // candidates are single-plane fetches (no averaging), up/down read the same plane one row apart and left/right
    // the same plane one pixel apart, so 9 + 8 row fetches serve all four.  Order and ties as qsad_best: up, down, left, right.
    __device__ __forceinline__ void hpel4_even( int omx, int omy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        const int s4 = pitch() >> 2, psz = (int)P->plane_size;
        u32 acc[4] = {0, 0, 0, 0};
        {   // up (omx, omy-2) is row 0..7, down (omx, omy+2) row 1..8 of the same plane
            const int y = omy - 2, pl = ( ( omx & 2 ) >> 1 ) | ( y & 2 );
            const RowSrc<LD64> src( ref0, ( y >> 2 ) * pitch() + ( omx >> 2 ) + pl * psz );
#pragma unroll
            for( int r = 0; r < 9; r++ )
            {
                const uint2 v = src.row( r*s4 );
                if( r < 8 ) acc[0] = __vsadu4( fenc[r].x, v.x ) + __vsadu4( fenc[r].y, v.y ) + acc[0];
                if( r > 0 ) acc[1] = __vsadu4( fenc[r-1].x, v.x ) + __vsadu4( fenc[r-1].y, v.y ) + acc[1];
            }
        }
        {   // left (omx-2, omy) starts at byte 0, right (omx+2, omy) at byte 1 of the same rows
            const int x = omx - 2, pl = ( ( x & 2 ) >> 1 ) | ( omy & 2 );
            const RowSrc<LD64> src( ref0, ( omy >> 2 ) * pitch() + ( x >> 2 ) + pl * psz );
#pragma unroll
            for( int r = 0; r < 8; r++ )
            {
                u32 t0, t1, t2;
                src.row12( r*s4, t0, t1, t2 );
                acc[2] = __vsadu4( fenc[r].x, t0 ) + __vsadu4( fenc[r].y, t1 ) + acc[2];
                acc[3] = __vsadu4( fenc[r].x, __funnelshift_r( t0, t1, 8 ) ) + __vsadu4( fenc[r].y, __funnelshift_r( t1, t2, 8 ) ) + acc[3];
            }
        }
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            const int x = omx + ( k == 2 ? -2 : k == 3 ? 2 : 0 ), y = omy + ( k == 0 ? -2 : k == 1 ? 2 : 0 );
            const int c = (int)acc[k] + mvcost( tab, m, x, y );
            if( c < bcost ) { bcost = c; best = k; }
        }
    }
    // The same diamond around ANY centre.  The two vertical candidates (omx, omy -+ 2) have the same quarter-pel phase, so
    // get_ref reads the same one or two planes at the same columns, one row apart: 9 rows of each plane serve both.  The
    // two horizontal candidates (omx -+ 2, omy) likewise read the same planes one pixel apart: 9 pixels per row serve both
    // (the byte average commutes with the byte shift, so the planes are averaged once, on three words).  34 plane-row
    // fetches instead of 64, in ONE code path for every lane -- a warp whose blocks disagree on the parity of their
    // centres no longer runs hpel4_even AND four generic get_ref candidates.
    __device__ __forceinline__ void hpel4_pairs( int omx, int omy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        static_assert( LD64 == 2 && PTXPRED, "hpel4_pairs is written for the 64+32-bit predicated row loader" );
        const int s4 = pitch() >> 2, psz = (int)P->plane_size;
        u32 acc[4] = {0, 0, 0, 0};
        {   // up (omx, omy-2): rows 0..7, down (omx, omy+2): rows 1..8
            const int y = omy - 2, qi = ( ( y & 3 ) << 2 ) + ( omx & 3 ), base = ( y >> 2 ) * pitch() + ( omx >> 2 );
            const int avg = ( qi & 5 ) != 0;
            const RowSrc<2> a( ref0, base + ( ( y & 3 ) == 3 ? pitch() : 0 ) + hpel_ref0( qi ) * psz ),
                            b( ref0, base + ( ( omx & 3 ) == 3 ) + hpel_ref1( qi ) * psz );
#pragma unroll
            for( int r = 0; r < 9; r++ )
            {
                uint2 v = a.row( r*s4 );
                const uint2 w = b.row_if( avg, r*s4, v );
                v.x = avg ? avg4( v.x, w.x ) : v.x; v.y = avg ? avg4( v.y, w.y ) : v.y;
                if( WEIGHTED && m.wdenom >= 0 ) { v.x = weight4( v.x, m.wscale, m.wdenom, m.woffset ); v.y = weight4( v.y, m.wscale, m.wdenom, m.woffset ); }
                if( r < 8 ) acc[0] = __vsadu4( fenc[r].x, v.x ) + __vsadu4( fenc[r].y, v.y ) + acc[0];
                if( r > 0 ) acc[1] = __vsadu4( fenc[r-1].x, v.x ) + __vsadu4( fenc[r-1].y, v.y ) + acc[1];
            }
        }
        {   // left (omx-2, omy): bytes 0..7, right (omx+2, omy): bytes 1..8 of the same rows
            const int x = omx - 2, qi = ( ( omy & 3 ) << 2 ) + ( x & 3 ), base = ( omy >> 2 ) * pitch() + ( x >> 2 );
            const int avg = ( qi & 5 ) != 0;
            const RowSrc<2> a( ref0, base + ( ( omy & 3 ) == 3 ? pitch() : 0 ) + hpel_ref0( qi ) * psz ),
                            b( ref0, base + ( ( x & 3 ) == 3 ) + hpel_ref1( qi ) * psz );
#pragma unroll
            for( int r = 0; r < 8; r++ )
            {
                u32 t0, t1, t2, u0, u1, u2;
                a.row12( r*s4, t0, t1, t2 );
                b.row12_if( avg, r*s4, u0, u1, u2 );
                t0 = avg ? avg4( t0, u0 ) : t0; t1 = avg ? avg4( t1, u1 ) : t1; t2 = avg ? avg4( t2, u2 ) : t2;
                if( WEIGHTED && m.wdenom >= 0 )
                {
                    t0 = weight4( t0, m.wscale, m.wdenom, m.woffset ); t1 = weight4( t1, m.wscale, m.wdenom, m.woffset );
                    t2 = weight4( t2, m.wscale, m.wdenom, m.woffset );
                }
                acc[2] = __vsadu4( fenc[r].x, t0 ) + __vsadu4( fenc[r].y, t1 ) + acc[2];
                acc[3] = __vsadu4( fenc[r].x, __funnelshift_r( t0, t1, 8 ) ) + __vsadu4( fenc[r].y, __funnelshift_r( t1, t2, 8 ) ) + acc[3];
            }
        }
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            const int x = omx + ( k == 2 ? -2 : k == 3 ? 2 : 0 ), y = omy + ( k == 0 ? -2 : k == 1 ? 2 : 0 );
            const int c = (int)acc[k] + mvcost( tab, m, x, y );
            if( c < bcost ) { bcost = c; best = k; }
        }
    }
    // The full hexagon (me.c:354-362) with shared row loads: the six points touch rows oy-2..oy+9 and columns
    // ox-2..ox+9, so each of the 12 rows is loaded once (4 words), aligned once to column ox-2, and the dx = -1/+1
    // candidates are immediate funnel shifts of it.  96 + 36 row loads become 48; costs are then compared in the
    // reference's order with strict '<'.
    __device__ __forceinline__ void hex6( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        const int s4 = pitch() >> 2, off = ( oy - 2 ) * pitch() + ox - 2;
        const u32 *q = (const u32 *)ref_w + ( off >> 2 ); const u32 sh = ( off & 3 ) * 8;
        u32 acc[6] = {0, 0, 0, 0, 0, 0};      // hex2[1..6]: (-2,0) (-1,2) (1,2) (2,0) (1,-2) (-1,-2)
#pragma unroll
        for( int r = 0; r < 12; r++ )
        {
            u32 a0, a1, a2;
            ld12w( q + r*s4, sh, a0, a1, a2 );
            if( r <= 7 || r >= 4 )
            {
                const u32 m0 = __funnelshift_r( a0, a1, 8 ),  m1 = __funnelshift_r( a1, a2, 8 );      // column ox-1
                const u32 p0 = __funnelshift_r( a0, a1, 24 ), p1 = __funnelshift_r( a1, a2, 24 );     // column ox+1
                if( r <= 7 )
                {   // dy = -2: block row r
                    acc[5] = __vsadu4( fenc[r].x, m0 ) + __vsadu4( fenc[r].y, m1 ) + acc[5];
                    acc[4] = __vsadu4( fenc[r].x, p0 ) + __vsadu4( fenc[r].y, p1 ) + acc[4];
                }
                if( r >= 4 )
                {   // dy = +2: block row r-4
                    acc[1] = __vsadu4( fenc[r-4].x, m0 ) + __vsadu4( fenc[r-4].y, m1 ) + acc[1];
                    acc[2] = __vsadu4( fenc[r-4].x, p0 ) + __vsadu4( fenc[r-4].y, p1 ) + acc[2];
                }
            }
            if( r >= 2 && r <= 9 )
            {   // dy = 0: block row r-2, columns ox-2 and ox+2
                acc[0] = __vsadu4( fenc[r-2].x, a0 ) + __vsadu4( fenc[r-2].y, a1 ) + acc[0];
                acc[3] = __vsadu4( fenc[r-2].x, a1 ) + __vsadu4( fenc[r-2].y, a2 ) + acc[3];
            }
        }
#pragma unroll
        for( int k = 0; k < 6; k++ )
        {
            const int x = ox + c_hex2[k+1][0], y = oy + c_hex2[k+1][1];
            const int c = (int)acc[k] + mvcost( tab, m, x*4, y*4 );
            if( c < bcost ) { bcost = c; best = k; }
        }
    }
    // The square refine (me.c:404-418): rows oy-1..oy+8, columns ox-1..ox+8, 10 shared row loads for 8 candidates
    __device__ __forceinline__ void square8( int ox, int oy, const CostTab &tab, const MeGeo &m, int &bcost, int &best ) const
    {
        const int s4 = pitch() >> 2, off = ( oy - 1 ) * pitch() + ox - 1;
        const u32 *q = (const u32 *)ref_w + ( off >> 2 ); const u32 sh = ( off & 3 ) * 8;
        u32 acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // square1[1..8]: (0,-1) (0,1) (-1,0) (1,0) (-1,-1) (-1,1) (1,-1) (1,1)
#pragma unroll
        for( int r = 0; r < 10; r++ )
        {
            u32 a0, a1, a2;
            ld12w( q + r*s4, sh, a0, a1, a2 );
            const u32 c0 = __funnelshift_r( a0, a1, 8 ),  c1 = __funnelshift_r( a1, a2, 8 );       // column ox
            const u32 p0 = __funnelshift_r( a0, a1, 16 ), p1 = __funnelshift_r( a1, a2, 16 );      // column ox+1
            if( r <= 7 )
            {   // dy = -1: block row r
                acc[0] = __vsadu4( fenc[r].x, c0 ) + __vsadu4( fenc[r].y, c1 ) + acc[0];
                acc[4] = __vsadu4( fenc[r].x, a0 ) + __vsadu4( fenc[r].y, a1 ) + acc[4];
                acc[6] = __vsadu4( fenc[r].x, p0 ) + __vsadu4( fenc[r].y, p1 ) + acc[6];
            }
            if( r >= 1 && r <= 8 )
            {   // dy = 0: block row r-1
                acc[2] = __vsadu4( fenc[r-1].x, a0 ) + __vsadu4( fenc[r-1].y, a1 ) + acc[2];
                acc[3] = __vsadu4( fenc[r-1].x, p0 ) + __vsadu4( fenc[r-1].y, p1 ) + acc[3];
            }
            if( r >= 2 )
            {   // dy = +1: block row r-2
                acc[1] = __vsadu4( fenc[r-2].x, c0 ) + __vsadu4( fenc[r-2].y, c1 ) + acc[1];
                acc[5] = __vsadu4( fenc[r-2].x, a0 ) + __vsadu4( fenc[r-2].y, a1 ) + acc[5];
                acc[7] = __vsadu4( fenc[r-2].x, p0 ) + __vsadu4( fenc[r-2].y, p1 ) + acc[7];
            }
        }
#pragma unroll
        for( int k = 0; k < 8; k++ )
        {
            const int x = ox + c_sq1[k+1][0], y = oy + c_sq1[k+1][1];
            const int c = (int)acc[k] + mvcost( tab, m, x*4, y*4 );
            if( c < bcost ) { bcost = c; best = k; }
        }
    }
    __device__ __forceinline__ int qsatd( const MeGeo &m, int mx, int my ) const
    {
        return satd8x8_t( fenc, QpelRows<WEIGHTED, PITCH, LD64, PTXPRED>( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) );
    }
    // SATD stage around a centre with even components (unweighted): the centre is one plane's rows K; a quarter-pel neighbour
    // (mx, my) is the rounding average of K with ONE other window -- whichever of get_ref's two sources is not the centre's.
    __device__ __forceinline__ void centre_rows( int omx, int omy, uint2 ( &K )[8] ) const
    {
        const int s4 = pitch() >> 2;
        const RowSrc<LD64> c( ref0, QpelRows<false, PITCH, LD64, PTXPRED>::off1( *P, omx, omy ) );
#pragma unroll
        for( int r = 0; r < 8; r++ ) K[r] = c.row( r*s4 );
    }
    __device__ __forceinline__ int qsatd_even( const uint2 ( &K )[8], int omx, int omy, int mx, int my ) const
    {
        typedef QpelRows<false, PITCH, LD64, PTXPRED> Q;
        const int s4 = pitch() >> 2, oc = Q::off1( *P, omx, omy ), o1 = Q::off1( *P, mx, my ), o2 = Q::off2( *P, mx, my );
        // the centre itself (mx == omx, my == omy): o1 == oc and the "other" window is the centre again; avg4( x, x ) == x
        const RowSrc<LD64> n( ref0, ( mx == omx && my == omy ) ? oc : o1 == oc ? o2 : o1 );
        return satd8x8_t( fenc, [&]( int r ) { const uint2 w = n.row( r*s4 ); return make_uint2( avg4( K[r].x, w.x ), avg4( K[r].y, w.y ) ); } );
    }
    __device__ __forceinline__ int bidir( int mv0x, int mv0y, int mv1x, int mv1y, int w ) const
    {
        if( !SATD )
        {
            const int s4 = pitch() >> 2, psz = (int)P->plane_size;
            const int i0 = ( ( mv0x & 2 ) >> 1 ) + ( mv0y & 2 ), i1 = ( ( mv1x & 2 ) >> 1 ) + ( mv1y & 2 );
            const int o0 = ( mv0x >> 2 ) + ( mv0y >> 2 ) * pitch() + i0*psz, o1 = ( mv1x >> 2 ) + ( mv1y >> 2 ) * pitch() + i1*psz;
            const RowSrc<LD64> q0( ref0, o0 ), q1( ref1, o1 );
#if X264CU_BIDIR_AHEAD >= 3
            uint2 A[8], Bv[8];
#pragma unroll
            for( int r = 0; r < 8; r++ ) { A[r] = q0.row( r*s4 ); Bv[r] = q1.row( r*s4 ); }
            return sad8x8_t( fenc, [&]( int r ) { return make_uint2( biavg4( A[r].x, Bv[r].x, w ), biavg4( A[r].y, Bv[r].y, w ) ); } );
#endif
            return sad8x8_t( fenc, [&]( int r ) { uint2 a = q0.row( r*s4 ), b = q1.row( r*s4 );
                                                  return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
        }
        QpelRows<false, PITCH, LD64, PTXPRED> r0( ref0, *P, mv0x, mv0y, 0, -1, 0 ), r1( ref1, *P, mv1x, mv1y, 0, -1, 0 );
#if X264CU_BIDIR_AHEAD
        // tuning: the whole list-0 prediction (8 rows, up to 32 loads) is requested before the transform starts, so that a warp
        // has more than one row of each plane in flight (the kernel is bound by the latency of its L2 -> L1 stream)
        uint2 A[8];
#pragma unroll
        for( int r = 0; r < 8; r++ ) A[r] = r0( r );
#if X264CU_BIDIR_AHEAD == 2
#pragma unroll
        for( int r = 0; r < 8; r++ ) asm volatile( "" : "+r"( A[r].x ), "+r"( A[r].y ) );
#endif
#if X264CU_BIDIR_AHEAD >= 3
        uint2 Bv[8];
#pragma unroll
        for( int r = 0; r < 8; r++ ) Bv[r] = r1( r );
        return satd8x8_t( fenc, [&]( int r ) { const uint2 a = A[r], b = Bv[r]; return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
#endif
        return satd8x8_t( fenc, [&]( int r ) { const uint2 a = A[r], b = r1( r ); return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
#else
        return satd8x8_t( fenc, [&]( int r ) { uint2 a = r0( r ), b = r1( r ); return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
#endif
    }
    __device__ __forceinline__ int bidir_zero( int w ) const
    {
        const u32 *q0 = (const u32 *)ref0, *q1 = (const u32 *)ref1; const int s4 = pitch() >> 2;
#if X264CU_BIDIR_AHEAD >= 4
        uint2 A[8], Bv[8];
#pragma unroll
        for( int r = 0; r < 8; r++ ) { A[r] = __ldg( (const uint2 *)( q0 + r*s4 ) ); Bv[r] = __ldg( (const uint2 *)( q1 + r*s4 ) ); }
        auto rows = [&]( int r ) { return make_uint2( biavg4( A[r].x, Bv[r].x, w ), biavg4( A[r].y, Bv[r].y, w ) ); };
#else
        auto rows = [&]( int r ) { uint2 a = __ldg( (const uint2 *)( q0 + r*s4 ) ), b = __ldg( (const uint2 *)( q1 + r*s4 ) );
                                   return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); };
#endif
        return SATD ? satd8x8_t( fenc, rows ) : sad8x8_t( fenc, rows );
    }
};


__device__ __forceinline__ bool skipped_border( const LaParams &P, int bx, int by )
{   // never evaluated by slicetype_slice_cost (slicetype.c:823-828)
    const bool border = bx == 0 || by == 0 || bx == P.mbw-1 || by == P.mbh-1;
    return border && !P.do_edges && P.mbw > 2 && P.mbh > 2;
}

// block owned by this lane under the warp-tile table; false: idle lane
__device__ __forceinline__ bool warp_tile_block( const uint32_t *__restrict__ tiles, int ntiles, const LaParams &P, int &bx, int &by )
{
    const int wt = blockIdx.x * ( blockDim.x >> 5 ) + ( threadIdx.x >> 5 ), lane = threadIdx.x & 31;
    if( wt >= ntiles ) return false;
    const uint32_t d = __ldg( tiles + wt );
    const int lw = d >> 24;
    bx = (int)( d & 0xfff ) + ( lane & ( ( 1 << lw ) - 1 ) );
    by = (int)( ( d >> 12 ) & 0xfff ) + ( lane >> lw );
    return bx < P.mbw && by < P.mbh;
}

static inline bool use_thread_mapping( const LaParams &P, int njobs )
{
    if( P.lanes_per_block == 1 ) return true;
    if( P.lanes_per_block == 8 ) return false;
    return (int64_t)njobs * P.nmb >= 4 * 8160;     // wide launches: enough blocks to fill the SMs one thread each
}

