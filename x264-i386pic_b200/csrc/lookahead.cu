// lookahead.cu -- sm_100a kernels of the lowres lookahead cost engine.
//
//   lowres_kernel     frame_init_lowres_core + border (mc.c:458-507, frame.c:535-555,627-631)
//   weight_kernel     x264_weight_scale_plane of the reference F plane (frame.c:825-840, mc.c:117-137)
//   intra_kernel      intra section of slicetype_mb_cost (slicetype.c:714-757)
//   search_kernel     x264_me_search_ref for PIXEL_8x8 (me.c:182-420,774-797,865-992) under the
//                     MV-predictor rule of slicetype.c:658-697, one pass of the dependency-free schedule
//   final_kernel      bidir candidates, intra/inter decision, AQ, packing, sums (slicetype.c:616-652,
//                     703-712,758-790,946-989)
//
// Thread mapping of the three block kernels: one 8-lane group per 8x8 block, lane r owning pixel
// row r (two 32-bit words).  SAD = VABSDIFF4.U8.ACC on the words, reduced with 3 xor-shuffles
// (two candidates per shuffle, packed 16:16); SATD = packed-lane Hadamard, horizontal stage in
// registers, vertical stage as a warp-shuffle butterfly over the rows of each 8x4 half.
// A CTA of 256 threads covers 32 consecutive blocks in raster order, so the 4 blocks of a warp are
// horizontal neighbours and their row loads fall in the same 128-byte lines.
#include "lookahead.cuh"

#define GROUPS_PER_CTA 32
#define CTA_THREADS (GROUPS_PER_CTA*8)

__constant__ uint8_t c_hpel_ref0[16] = {0,1,1,1,0,1,1,1,2,3,3,3,0,1,1,1};   // tables.c:183
__constant__ uint8_t c_hpel_ref1[16] = {0,0,1,0,2,2,3,2,2,2,3,2,2,2,3,2};   // tables.c:184

// ------------------------------------------------------------------------------------------------
// lowres
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lum( const uint8_t *luma, int ls, int W, int H, int x, int y )
{   // mod-16 padding (frame.c:640-668) and the +1 row/column dup (mc.c:465-468) are pure replication
    return luma[(size_t)min( y, H-1 ) * ls + min( x, W-1 )];
}
#define FILT(a,b,c,d) ((((((a)+(b)+1)>>1)+(((c)+(d)+1)>>1)+1)>>1))

__global__ void __launch_bounds__(128) lowres_kernel( const uint8_t *__restrict__ luma, int ls, int W, int H,
                                                      uint8_t *__restrict__ planes, LaParams P )
{
    const int pw4 = ( P.lw + 2*X264CU_PAD ) >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y;
    if( g >= pw4 ) return;
    const int y = min( max( py - X264CU_PAD, 0 ), P.lh - 1 );
    const int x0 = g*4 - X264CU_PAD;
    u32 o[4];
    const uint8_t *r0 = luma + (size_t)min( 2*y,   H-1 ) * ls;
    const uint8_t *r1 = luma + (size_t)min( 2*y+1, H-1 ) * ls;
    const uint8_t *r2 = luma + (size_t)min( 2*y+2, H-1 ) * ls;
    const bool fast = x0 >= 0 && x0 + 4 <= P.lw && 2*x0 + 8 <= W-1 && ( ( (uintptr_t)luma | (uintptr_t)ls ) & 7 ) == 0;
    if( fast )
    {   // 9 input pixels per row: two aligned words + one byte
        const int c = 2*x0;
        uint2 a = *(const uint2 *)( r0 + c ), b = *(const uint2 *)( r1 + c ), d = *(const uint2 *)( r2 + c );
        u32 a8 = r0[c+8], b8 = r1[c+8], d8 = r2[c+8];
        u32 v0 = avg4( a.x, b.x ), v1 = avg4( a.y, b.y ), v8 = ( a8 + b8 + 1 ) >> 1;   // rows 2y,2y+1
        u32 w0 = avg4( b.x, d.x ), w1 = avg4( b.y, d.y ), w8 = ( b8 + d8 + 1 ) >> 1;   // rows 2y+1,2y+2
        u32 ve = __byte_perm( v0, v1, 0x6420 ), vo = __byte_perm( v0, v1, 0x7531 ), ve2 = ( ve >> 8 ) | ( v8 << 24 );
        u32 we = __byte_perm( w0, w1, 0x6420 ), wo = __byte_perm( w0, w1, 0x7531 ), we2 = ( we >> 8 ) | ( w8 << 24 );
        o[0] = avg4( ve, vo ); o[1] = avg4( vo, ve2 ); o[2] = avg4( we, wo ); o[3] = avg4( wo, we2 );
    }
    else
    {
        o[0] = o[1] = o[2] = o[3] = 0;
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            int x = min( max( x0 + k, 0 ), P.lw - 1 );
            int p00 = lum( luma, ls, W, H, 2*x, 2*y ),   p10 = lum( luma, ls, W, H, 2*x+1, 2*y ),   p20 = lum( luma, ls, W, H, 2*x+2, 2*y );
            int p01 = lum( luma, ls, W, H, 2*x, 2*y+1 ), p11 = lum( luma, ls, W, H, 2*x+1, 2*y+1 ), p21 = lum( luma, ls, W, H, 2*x+2, 2*y+1 );
            int p02 = lum( luma, ls, W, H, 2*x, 2*y+2 ), p12 = lum( luma, ls, W, H, 2*x+1, 2*y+2 ), p22 = lum( luma, ls, W, H, 2*x+2, 2*y+2 );
            o[0] |= (u32)FILT( p00, p01, p10, p11 ) << ( 8*k );
            o[1] |= (u32)FILT( p10, p11, p20, p21 ) << ( 8*k );
            o[2] |= (u32)FILT( p01, p02, p11, p12 ) << ( 8*k );
            o[3] |= (u32)FILT( p11, p12, p21, p22 ) << ( 8*k );
        }
    }
    size_t off = (size_t)py * P.stride + (size_t)g*4;      // planes[] points at the padded origin
#pragma unroll
    for( int k = 0; k < 4; k++ )
        *(u32 *)( planes + k*P.plane_size + off ) = o[k];
}

void launch_lowres( const uint8_t *luma, int stride, int width, int height, uint8_t *planes, const LaParams &P, cudaStream_t st )
{
    int pw4 = ( P.lw + 2*X264CU_PAD ) >> 2;
    dim3 grid( ( pw4 + 127 ) / 128, P.lh + 2*X264CU_PAD );
    lowres_kernel<<<grid, 128, 0, st>>>( luma, stride, width, height, planes, P );
}

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

__global__ void weight_kernel( const u32 *__restrict__ src, u32 *__restrict__ dst, int64_t nwords, int scale, int denom, int offset )
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( i < nwords ) dst[i] = weight4( src[i], scale, denom, offset );
}

void launch_weight_plane( const uint8_t *src, uint8_t *dst, int64_t bytes, int scale, int denom, int offset, cudaStream_t st )
{
    int64_t nwords = bytes >> 2;
    weight_kernel<<<(unsigned)( ( nwords + 255 ) / 256 ), 256, 0, st>>>( (const u32 *)src, (u32 *)dst, nwords, scale, denom, offset );
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
    const u32 w2 = 64 - w;
    u32 ae = a & 0x00ff00ffu, ao = ( a >> 8 ) & 0x00ff00ffu, be = b & 0x00ff00ffu, bo = ( b >> 8 ) & 0x00ff00ffu;
    u32 re = ( ( ae * w + be * w2 + 0x00200020u ) >> 6 ) & 0x00ff00ffu;
    u32 ro = ( ( ao * w + bo * w2 + 0x00200020u ) << 2 ) & 0xff00ff00u;
    return re | ro;
}

// ------------------------------------------------------------------------------------------------
// intra
// ------------------------------------------------------------------------------------------------
#define F1(a,b)   (((a)+(b)+1)>>1)
#define F2(a,b,c) (((a)+2*(b)+(c)+2)>>2)

__global__ void __launch_bounds__(CTA_THREADS) intra_kernel( const uint8_t *__restrict__ planeF, const uint16_t *__restrict__ inv_qscale,
                                                             uint16_t *__restrict__ intra, int32_t *row_satds, int32_t *stats, LaParams P )
{
    // per group: raw[0]=tl, raw[1..16]=t0..t15, raw[17..24]=l0..l7 ; v[0..7]=l7..l0 filtered, v[8]=lt, v[9..24]=t0..t15 filtered
    __shared__ uint8_t s_raw[GROUPS_PER_CTA][28];
    __shared__ uint8_t s_v[GROUPS_PER_CTA][28];
    const int gi = threadIdx.x >> 3;
    Grp g; g.r = threadIdx.x & 7; g.mask = 0xffu << ( threadIdx.x & 24 );
    const int blk = blockIdx.x * GROUPS_PER_CTA + gi;
    if( blk >= P.nmb ) return;
    const int bx = blk % P.mbw, by = blk / P.mbw;
    const bool border = bx == 0 || by == 0 || bx == P.mbw-1 || by == P.mbh-1;
    const bool tiny = P.mbw <= 2 || P.mbh <= 2;
    if( border && !P.do_edges && !tiny )
    {   // never evaluated by slicetype_slice_cost (slicetype.c:823-828): keeps frame_new's 0xFFFF
        if( g.r == 0 ) intra[blk] = 0xFFFF;
        return;
    }
    const uint8_t *src = planeF + (size_t)( by*8 ) * P.stride + bx*8;
    uint8_t *raw = s_raw[gi], *v = s_v[gi];
    const int r = g.r;
    // neighbours (slicetype.c:724-727): 16 above (+ top-left), 8 to the left
    raw[1+r] = src[r - P.stride];
    raw[9+r] = src[r + 8 - P.stride];
    raw[17+r] = src[-1 + r*P.stride];
    if( r == 0 ) raw[0] = src[-1 - P.stride];
    __syncwarp( g.mask );
#define T(i) ((int)raw[1+(i)])
#define L(i) ((int)raw[17+(i)])
#define TL   ((int)raw[0])
    // predict_8x8_filter_c with all neighbours (predict.c:632-676)
    {
        int fl = r == 0 ? ( TL + 2*L(0) + L(1) + 2 ) >> 2 : r == 7 ? ( L(6) + 3*L(7) + 2 ) >> 2 : F2( L(r-1), L(r), L(r+1) );
        v[7-r] = fl;
        int ft = r == 0 ? ( TL + 2*T(0) + T(1) + 2 ) >> 2 : F2( T(r-1), T(r), T(r+1) );
        v[9+r] = ft;
        int ft2 = r == 7 ? ( T(14) + 3*T(15) + 2 ) >> 2 : F2( T(r+7), T(r+8), T(r+9) );
        v[17+r] = ft2;
        if( r == 0 ) v[8] = ( T(0) + 2*TL + L(0) + 2 ) >> 2;
    }
    __syncwarp( g.mask );
    uint2 fenc = *(const uint2 *)( src + (size_t)r * P.stride );
    int best;
    uint2 p;
#define PACK8( expr ) { u32 lo_ = 0, hi_ = 0; _Pragma("unroll") for( int x = 0; x < 4; x++ ) { lo_ |= (u32)(expr) << (8*x); } \
                        _Pragma("unroll") for( int x = 4; x < 8; x++ ) { hi_ |= (u32)(expr) << (8*(x-4)); } p = make_uint2( lo_, hi_ ); }
#define TRY() { int c_ = cmp_rows( fenc, p, g, P.b_satd ); best = min( best, c_ ); }
    {   // 8x8c DC (predict.c:221-257)
        int s0 = T(0)+T(1)+T(2)+T(3), s1 = T(4)+T(5)+T(6)+T(7), s2 = L(0)+L(1)+L(2)+L(3), s3 = L(4)+L(5)+L(6)+L(7);
        int dl = r < 4 ? ( s0 + s2 + 4 ) >> 3 : ( s3 + 2 ) >> 2;
        int dr = r < 4 ? ( s1 + 2 ) >> 2 : ( s1 + s3 + 4 ) >> 3;
        p = make_uint2( dl * 0x01010101u, dr * 0x01010101u );
        best = cmp_rows( fenc, p, g, P.b_satd );
    }
    p = make_uint2( L(r) * 0x01010101u, L(r) * 0x01010101u ); TRY()                       // 8x8c H
    PACK8( T(x) ) TRY()                                                                     // 8x8c V
    if( P.b_satd )
    {
        {   // 8x8c plane (predict.c:282-304)
            int H = 0, V = 0;
#pragma unroll
            for( int i = 0; i < 4; i++ )
            {
                H += ( i+1 ) * ( T(4+i) - ( i == 3 ? TL : T(2-i) ) );
                V += ( i+1 ) * ( L(4+i) - ( i == 3 ? TL : L(2-i) ) );
            }
            int a = 16 * ( L(7) + T(7) ), b = ( 17*H + 16 ) >> 5, c = ( 17*V + 16 ) >> 5;
            int i00 = a - 3*b - 3*c + 16 + r*c;
            PACK8( min( max( ( i00 + x*b ) >> 5, 0 ), 255 ) ) TRY()
        }
#define V_(i) ((int)v[(i)])
#define FT(i) ((int)v[9+(i)])
#define FL(i) ((int)v[7-(i)])
        const int y = r;
        // DDL (predict.c:741-760)
        PACK8( ( x+y < 14 ? F2( FT(x+y), FT(x+y+1), FT(x+y+2) ) : F2( FT(14), FT(15), FT(15) ) ) ) TRY()
        // DDR (:761-781): centre tap v[8+x-y]
        PACK8( F2( V_(7+x-y), V_(8+x-y), V_(9+x-y) ) ) TRY()
        // VR (:783-811)
        PACK8( ( 2*x-y >= 0 ? ( ( (2*x-y) & 1 ) ? F2( V_(7+x-(y>>1)), V_(8+x-(y>>1)), V_(9+x-(y>>1)) ) : F1( V_(8+x-(y>>1)), V_(9+x-(y>>1)) ) )
                 : 2*x-y == -1 ? F2( V_(7), V_(8), V_(9) )
                 : F2( V_(8-(y-2*x-1)-1), V_(8-(y-2*x-1)), V_(8-(y-2*x-1)+1) ) ) ) TRY()
        // HD (:812-837)
        PACK8( ( 2*y-x >= 0 ? ( ( (2*y-x) & 1 ) ? F2( V_(7-(y-(x>>1))), V_(8-(y-(x>>1))), V_(9-(y-(x>>1))) ) : F1( V_(7-(y-(x>>1))), V_(8-(y-(x>>1))) ) )
                 : 2*y-x == -1 ? F2( V_(7), V_(8), V_(9) )
                 : F2( V_(8+(x-2*y-1)-1), V_(8+(x-2*y-1)), V_(8+(x-2*y-1)+1) ) ) ) TRY()
        // VL (:838-864)
        PACK8( ( ( y & 1 ) ? F2( FT(x+(y>>1)), FT(x+(y>>1)+1), FT(x+(y>>1)+2) ) : F1( FT(x+(y>>1)), FT(x+(y>>1)+1) ) ) ) TRY()
        // HU (:865-887)
        PACK8( ( x+2*y > 13 ? FL(7) : x+2*y == 13 ? F2( FL(6), FL(7), FL(7) )
                 : ( (x+2*y) & 1 ) ? F2( FL((x+2*y)>>1), FL(((x+2*y)>>1)+1), FL(((x+2*y)>>1)+2) ) : F1( FL((x+2*y)>>1), FL(((x+2*y)>>1)+1) ) ) ) TRY()
    }
    if( g.r == 0 )
    {
        int icost = best + 5 + 4;                 // + 5*lambda + lowres_penalty (slicetype.c:720,746)
        intra[blk] = (uint16_t)icost;
        int icost_aq = P.aq ? ( icost * inv_qscale[blk] + 128 ) >> 8 : icost;
        atomicAdd( &row_satds[by], icost_aq );
        if( !border || tiny )
        {
            atomicAdd( &stats[0], icost );
            atomicAdd( &stats[1], icost_aq );
        }
    }
}

void launch_intra( const uint8_t *planeF, const uint16_t *inv_qscale, uint16_t *intra, int32_t *row_satds, int32_t *stats,
                   const LaParams &P, cudaStream_t st )
{
    intra_kernel<<<( P.nmb + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, CTA_THREADS, 0, st>>>( planeF, inv_qscale, intra, row_satds, stats, P );
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

// cost_mv table (analyse.c:153-157): the first S_TAB_N entries (|mvd| < 128 px) are staged in shared memory per CTA,
// the rare larger deltas read the global table through L1 (ncu r01: copying all 8*mv_range+1 entries per CTA was
// 14 % of the search kernel's stall samples)
#define S_TAB_N 512
struct CostTab
{
    const uint16_t *s, *g;
    __device__ __forceinline__ int operator[]( int d ) const { return d < S_TAB_N ? (int)s[d] : (int)__ldg( g + d ); }
};
__device__ __forceinline__ CostTab stage_cost_tab( uint16_t *s_tab, const LaParams &P )
{
    for( int i = threadIdx.x; i < S_TAB_N && i <= P.cost_mv_n; i += blockDim.x ) s_tab[i] = __ldg( P.cost_mv + i );
    __syncthreads();
    CostTab t; t.s = s_tab; t.g = P.cost_mv;
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

// 8 pixels at word pointer q (aligned) shifted by sh bits: the row loader of ThreadBE
__device__ __forceinline__ uint2 ld8w( const u32 *q, u32 sh )
{
    u32 w0 = __ldg( q ), w1 = __ldg( q + 1 ), w2 = __ldg( q + 2 );
    return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
}

__constant__ int8_t c_hex2[8][2] = {{-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0}};   // me.c:55
__constant__ int8_t c_sq1[9][2] = {{0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1}}; // me.c:56

// ---- 8 lanes per block -------------------------------------------------------------------------------
struct GroupBE
{
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
        u32 sum = 0;
#pragma unroll
        for( int i = 0; i < 4; i++ )
        {
            u32 a0, a1, a2, a3;
            HADAMARD4( a0, a1, a2, a3, t[0][i], t[1][i], t[2][i], t[3][i] );
            sum += abs2( a0 ) + abs2( a1 ) + abs2( a2 ) + abs2( a3 );
        }
        total += (int)( ( ( sum & 0xffff ) + ( sum >> 16 ) ) >> 1 );
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
template<bool WEIGHTED, int PITCH>
struct QpelRows
{
    const u32 *q1, *q2;      // word pointers of the two source planes at row 0
    u32 sh1, sh2;            // bit shifts
    int stride4;             // plane pitch in words
    bool avg;
    int wscale, wdenom, woffset;
    __device__ __forceinline__ QpelRows( const uint8_t *ref, const LaParams &P, int mvx, int mvy, int ws, int wd, int wo )
    {   // ref: 8-byte aligned block origin in the F plane; plane_size and the pitch are multiples of 4
        const int pitch = PITCH ? PITCH : P.stride, psz = (int)P.plane_size;
        const int qidx = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
        const int off = ( mvy >> 2 ) * pitch + ( mvx >> 2 );
        const int o1 = off + ( ( mvy & 3 ) == 3 ? pitch : 0 ) + c_hpel_ref0[qidx] * psz;
        const int o2 = off + ( ( mvx & 3 ) == 3 ) + c_hpel_ref1[qidx] * psz;
        q1 = (const u32 *)ref + ( o1 >> 2 ); sh1 = ( o1 & 3 ) * 8;
        q2 = (const u32 *)ref + ( o2 >> 2 ); sh2 = ( o2 & 3 ) * 8;
        stride4 = pitch >> 2;
        avg = ( qidx & 5 ) != 0;
        wscale = ws; wdenom = wd; woffset = wo;
    }
    __device__ __forceinline__ uint2 operator()( int r ) const
    {
        const int s4 = PITCH ? PITCH >> 2 : stride4;
        uint2 v = ld8w( q1 + r*s4, sh1 );
        if( avg )
        {
            uint2 w = ld8w( q2 + r*s4, sh2 );
            v.x = avg4( v.x, w.x ); v.y = avg4( v.y, w.y );
        }
        if( WEIGHTED && wdenom >= 0 ) { v.x = weight4( v.x, wscale, wdenom, woffset ); v.y = weight4( v.y, wscale, wdenom, woffset ); }
        return v;
    }
};

template<bool SATD, bool WEIGHTED, int PITCH>
struct ThreadBE
{
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
        const int s4 = pitch() >> 2, off = my * pitch() + mx;
        const u32 *q = (const u32 *)ref_w + ( off >> 2 ); const u32 sh = ( off & 3 ) * 8;
        return sad8x8_t( fenc, [&]( int r ) { return ld8w( q + r*s4, sh ); } );
    }
    __device__ __forceinline__ int qsad1( const MeGeo &m, int mx, int my ) const
    { return sad8x8_t( fenc, QpelRows<WEIGHTED, PITCH>( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) ); }
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
        return satd8x8_t( fenc, QpelRows<WEIGHTED, PITCH>( ref0, *P, mx, my, m.wscale, m.wdenom, m.woffset ) );
    }
    __device__ __forceinline__ int bidir( int mv0x, int mv0y, int mv1x, int mv1y, int w ) const
    {
        if( !SATD )
        {
            const int s4 = pitch() >> 2, psz = (int)P->plane_size;
            const int i0 = ( ( mv0x & 2 ) >> 1 ) + ( mv0y & 2 ), i1 = ( ( mv1x & 2 ) >> 1 ) + ( mv1y & 2 );
            const int o0 = ( mv0x >> 2 ) + ( mv0y >> 2 ) * pitch() + i0*psz, o1 = ( mv1x >> 2 ) + ( mv1y >> 2 ) * pitch() + i1*psz;
            const u32 *q0 = (const u32 *)ref0 + ( o0 >> 2 ), *q1 = (const u32 *)ref1 + ( o1 >> 2 );
            const u32 h0 = ( o0 & 3 ) * 8, h1 = ( o1 & 3 ) * 8;
            return sad8x8_t( fenc, [&]( int r ) { uint2 a = ld8w( q0 + r*s4, h0 ), b = ld8w( q1 + r*s4, h1 );
                                                  return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
        }
        QpelRows<false, PITCH> r0( ref0, *P, mv0x, mv0y, 0, -1, 0 ), r1( ref1, *P, mv1x, mv1y, 0, -1, 0 );
        return satd8x8_t( fenc, [&]( int r ) { uint2 a = r0( r ), b = r1( r ); return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
    }
    __device__ __forceinline__ int bidir_zero( int w ) const
    {
        const u32 *q0 = (const u32 *)ref0, *q1 = (const u32 *)ref1; const int s4 = pitch() >> 2;
        auto rows = [&]( int r ) { uint2 a = __ldg( (const uint2 *)( q0 + r*s4 ) ), b = __ldg( (const uint2 *)( q1 + r*s4 ) );
                                   return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); };
        return SATD ? satd8x8_t( fenc, rows ) : sad8x8_t( fenc, rows );
    }
};


#define FPEL(v) (((v)+2)>>2)

// refine_subpel (me.c:865-992), p_halfpel_thresh == NULL, b_refine_qpel == 0, no chroma
template<class BE>
__device__ __forceinline__ void refine_subpel( const BE &be, const MeGeo &m, const LaParams &P, const CostTab &tab,
                                               int &bmx, int &bmy, int &bcost, int hpel_iters, int qpel_iters )
{
    if( hpel_iters )
    {
        if( !be.satd() )
        {   // la subme 2 (< 3): try the subpel component of the predictor (me.c:889-895)
            int mx = iclip3( m.mvpx, m.sminx + 2, m.smaxx - 2 );
            int my = iclip3( m.mvpy, m.sminy + 2, m.smaxy - 2 );
            if( ( mx - bmx ) | ( my - bmy ) )
            {
                int cost = be.qsad1( m, mx, my ) + mvcost( tab, m, mx, my );
                if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; }
            }
        }
        for( int i = hpel_iters; i > 0; i-- )
        {   // me.c:898-921, tie order: up, down, left, right
            const int omx = bmx, omy = bmy;
            int best = -1;
            be.qsad_best( 4, [&]( int k, int &x, int &y ) { x = omx + ( k == 2 ? -2 : k == 3 ? 2 : 0 ); y = omy + ( k == 0 ? -2 : k == 1 ? 2 : 0 ); },
                          tab, m, bcost, best );
            if( best < 0 ) break;
            bmx = omx + ( best == 2 ? -2 : best == 3 ? 2 : 0 );
            bmy = omy + ( best == 0 ? -2 : best == 1 ? 2 : 0 );
        }
    }
    if( be.satd() )   // mbcmp != fpelcmp: re-measure the winner with SATD (me.c:925-929)
        bcost = be.qsatd( m, bmx, bmy ) + mvcost( tab, m, bmx, bmy );
    int bdir = -1;
    for( int i = qpel_iters; i > 0; i-- )
    {   // me.c:947-963
        if( bmy <= m.sminy || bmy >= m.smaxy || bmx <= m.sminx || bmx >= m.smaxx ) break;
        const int odir = bdir, omx = bmx, omy = bmy;
#pragma unroll 1
        for( int k = 0; k < 4; k++ )
            if( ( k ^ 1 ) != odir )
            {
                int mx = omx + ( k == 2 ? -1 : k == 3 ? 1 : 0 ), my = omy + ( k == 0 ? -1 : k == 1 ? 1 : 0 );
                int cost = be.qsatd( m, mx, my ) + mvcost( tab, m, mx, my );
                if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; bdir = k; }
            }
        if( bmx == omx && bmy == omy ) break;
    }
}

// x264_me_search_ref (me.c:182-420,774-797) for one block; result mv (qpel) and cost
template<class BE>
__device__ __forceinline__ void me_search( const BE &be, const MeGeo &m, const LaParams &P, const CostTab &tab,
                                           const int (*mvc)[2], int i_mvc, int &out_mx, int &out_my, int &out_cost )
{
    int bmx, bmy, bcost = X264CU_COST_MAX, bpred_cost = X264CU_COST_MAX, bpred_mx = 0, bpred_my = 0, pmvx, pmvy;
#define COST_MV( mx_, my_ ) { int c_ = be.fpel1( mx_, my_ ) + mvcost( tab, m, (mx_)*4, (my_)*4 ); \
                              if( c_ < bcost ) { bcost = c_; bmx = mx_; bmy = my_; } }
    if( be.satd() )
    {   // la subme 4 (>= 3): subpel predictors (me.c:216-275)
        bpred_mx = iclip3( m.mvpx, 4*m.fminx, 4*m.fmaxx );
        bpred_my = iclip3( m.mvpy, 4*m.fminy, 4*m.fmaxy );
        pmvx = bpred_mx; pmvy = bpred_my;
        bpred_cost = be.qsad1( m, bpred_mx, bpred_my ) + mvcost( tab, m, bpred_mx, bpred_my );
        const int pmv_cost = bpred_cost;
#pragma unroll 1
        for( int i = 0; i < i_mvc; i++ )
        {   // x264_predictor_clip (common.h:790-805): drop zero and == pmv, clip the rest
            int mx = mvc[i][0], my = mvc[i][1];
            if( ( !mx && !my ) || ( mx == pmvx && my == pmvy ) ) continue;
            mx = iclip3( mx, 4*m.fminx, 4*m.fmaxx ); my = iclip3( my, 4*m.fminy, 4*m.fmaxy );
            int cost = be.qsad1( m, mx, my ) + mvcost( tab, m, mx, my );
            if( cost < bpred_cost ) { bpred_cost = cost; bpred_mx = mx; bpred_my = my; }
        }
        bmx = FPEL( bpred_mx ); bmy = FPEL( bpred_my );
        // me.c:258-274 with the two possible fullpel probes folded into one call site
        const bool probe_start = ( bpred_mx & 3 ) | ( bpred_my & 3 );
        if( !probe_start ) bcost = bpred_cost;
        const bool probe_zero = ( pmvx | pmvy ) && ( bmx | bmy );
        const int sx = bmx, sy = bmy;
#pragma unroll 1
        for( int t = 0; t < 2; t++ )
        {
            if( t == 0 ? !probe_start : !probe_zero ) continue;
            const int x = t == 0 ? sx : 0, y = t == 0 ? sy : 0;
            COST_MV( x, y )
        }
        if( !( pmvx | pmvy ) && pmv_cost < bcost ) { bcost = pmv_cost; bmx = 0; bmy = 0; }
    }
    else
    {   // la subme 2: fullpel predictors (me.c:278-317)
        bmx = iclip3( FPEL( m.mvpx ), m.fminx, m.fmaxx );
        bmy = iclip3( FPEL( m.mvpy ), m.fminy, m.fmaxy );
        pmvx = bmx; pmvy = bmy;
        int sx = bmx, sy = bmy;
        // the start point (no mv cost, me.c:291), the rounded candidates (:293-313) and zero (:316-317) share one call site
#pragma unroll 1
        for( int i = -1; i <= i_mvc; i++ )
        {
            int mx, my, add;
            if( i < 0 ) { mx = bmx; my = bmy; add = 0; }
            else if( i < i_mvc )
            {   // x264_predictor_roundclip (common.h:774-788)
                mx = ( mvc[i][0] + 2 ) >> 2; my = ( mvc[i][1] + 2 ) >> 2;
                if( ( !mx && !my ) || ( mx == pmvx && my == pmvy ) ) continue;
                mx = iclip3( mx, m.fminx, m.fmaxx ); my = iclip3( my, m.fminy, m.fmaxy );
                add = mvcost( tab, m, mx*4, my*4 );
            }
            else
            {
                if( !( pmvx | pmvy ) ) continue;
                mx = 0; my = 0; add = mvcost( tab, m, 0, 0 );
            }
            int cost = be.fpel1( mx, my ) + add;
            if( cost < bcost ) { bcost = cost; sx = mx; sy = my; }
        }
        bmx = sx; bmy = sy;
    }
#undef COST_MV

    if( P.me_method == 0 )
    {   // DIA (me.c:322-342): tie order up, down, left, right
        int i = P.me_range;
        do
        {
            const int ox = bmx, oy = bmy;
            int best = -1;
            be.fpel_best( 4, [&]( int k, int &x, int &y ) { x = ox + ( k == 2 ? -1 : k == 3 ? 1 : 0 ); y = oy + ( k == 0 ? -1 : k == 1 ? 1 : 0 ); },
                          tab, m, bcost, best );
            if( best < 0 ) break;
            bmx = ox + ( best == 2 ? -1 : best == 3 ? 1 : 0 );
            bmy = oy + ( best == 0 ? -1 : best == 1 ? 1 : 0 );
        } while( --i && in_range( m, bmx, bmy ) );
    }
    else
    {   // HEX (me.c:344-420): the full hexagon, then half hexagons that do not overlap the previous one
        int best = -1;
        be.hex6( bmx, bmy, tab, m, bcost, best );
        if( best >= 0 )
        {
            int dir = best;
            bmx += c_hex2[dir+1][0]; bmy += c_hex2[dir+1][1];
            for( int i = ( P.me_range >> 1 ) - 1; i > 0 && in_range( m, bmx, bmy ); i-- )
            {
                const int ox = bmx, oy = bmy, base = dir;
                best = -1;
                be.fpel_best( 3, [&]( int k, int &x, int &y ) { x = ox + c_hex2[base+k][0]; y = oy + c_hex2[base+k][1]; }, tab, m, bcost, best );
                if( best < 0 ) break;
                dir += best - 1;
                dir = dir < 0 ? dir + 6 : dir >= 6 ? dir - 6 : dir;      // mod6m1 (me.c:53)
                bmx = ox + c_hex2[dir+1][0]; bmy = oy + c_hex2[dir+1][1];
            }
        }
        // square refine (me.c:404-418)
        {
            const int ox = bmx, oy = bmy;
            int bs = -1;
            be.square8( ox, oy, tab, m, bcost, bs );
            bmx = ox + c_sq1[bs+1][0]; bmy = oy + c_sq1[bs+1][1];
        }
    }

    // -> qpel mv (me.c:774-797)
    int mx, my, cost;
    if( !be.satd() )
    {
        cost = bcost;
        if( bmx == pmvx && bmy == pmvy ) cost += mvcost( tab, m, bmx*4, bmy*4 );
        mx = bmx*4; my = bmy*4;
        refine_subpel( be, m, P, tab, mx, my, cost, 1, 0 );      // subpel_iterations[2] = {0,1,1,0}
    }
    else
    {
        if( bpred_cost < bcost ) { mx = bpred_mx; my = bpred_my; }
        else                     { mx = bmx*4;    my = bmy*4; }
        cost = min( bpred_cost, bcost );
        refine_subpel( be, m, P, tab, mx, my, cost, 1, 1 );      // subpel_iterations[4] = {0,2,1,1}
    }
    out_mx = mx; out_my = my; out_cost = cost;
}

// one block of one search pass (slicetype.c:654-702)
template<class BE>
__device__ __forceinline__ void search_block( const BE &be, const SearchJob &J, const LaParams &P, const CostTab &tab,
                                              int bx, int by, int blk )
{
    MeGeo m;
    m.wscale = J.wscale; m.wdenom = J.wdenom; m.woffset = J.woffset;
    me_limits( m, P, bx, by );
    // MV predictor from the previous pass's field (slicetype.c:658-680): right, below, below-left, below-right
    int mvc[4][2] = {{0,0},{0,0},{0,0},{0,0}}, i_mvc = 0;
    if( J.mv_in )
    {
        const int *f = (const int *)J.mv_in + blk;
#define MVC( o ) { int v_ = __ldg( f + (o) ); mvc[i_mvc][0] = (int16_t)( v_ & 0xffff ); mvc[i_mvc][1] = v_ >> 16; i_mvc++; }
        if( bx < P.mbw-1 ) MVC( 1 )
        if( by < P.mbh-1 )
        {
            MVC( P.mbw )
            if( bx > 0 ) MVC( P.mbw-1 )
            if( bx < P.mbw-1 ) MVC( P.mbw+1 )
        }
#undef MVC
    }
    if( i_mvc <= 1 ) { m.mvpx = mvc[0][0]; m.mvpy = mvc[0][1]; }
    else
    {
        m.mvpx = imedian( mvc[0][0], mvc[1][0], mvc[2][0] );
        m.mvpy = imedian( mvc[0][1], mvc[1][1], mvc[2][1] );
    }
    int mx = 0, my = 0, cost = 0;
    bool skip = false;
    if( !( m.mvpx | m.mvpy ) )
    {   // zero-predictor shortcut (slicetype.c:684-692): unweighted F plane at mv 0
        cost = be.zero_cmp();
        skip = cost < 64;
    }
    if( !skip )
    {
        me_search( be, m, P, tab, mvc, i_mvc, mx, my, cost );
        cost -= tab[0];                                // slicetype.c:695
        if( mx | my ) cost += 5;                       // 5 * lambda (:696-697)
    }
    if( be.leader() )
    {
        ( (int *)J.mv_out )[blk] = ( mx & 0xffff ) | ( my << 16 );
        J.cost_out[blk] = cost;
    }
}

__device__ __forceinline__ bool skipped_border( const LaParams &P, int bx, int by )
{   // never evaluated by slicetype_slice_cost (slicetype.c:823-828)
    const bool border = bx == 0 || by == 0 || bx == P.mbw-1 || by == P.mbh-1;
    return border && !P.do_edges && P.mbw > 2 && P.mbh > 2;
}

__global__ void __launch_bounds__(CTA_THREADS) search_kernel_g( const SearchJob *__restrict__ jobs, LaParams P )
{
    __shared__ uint16_t s_tab[S_TAB_N];
    const CostTab tab = stage_cost_tab( s_tab, P );
    const SearchJob J = jobs[blockIdx.y];
    GroupBE be; be.g.r = threadIdx.x & 7; be.g.mask = 0xffu << ( threadIdx.x & 24 ); be.P = &P;
    const int blk = blockIdx.x * GROUPS_PER_CTA + ( threadIdx.x >> 3 );
    if( blk >= P.nmb ) return;
    const int bx = blk % P.mbw, by = blk / P.mbw;
    if( skipped_border( P, bx, by ) )
    {   // the field keeps frame_new's zeros (frame.c:283-285)
        if( be.g.r == 0 ) { ( (int *)J.mv_out )[blk] = 0; J.cost_out[blk] = 0; }
        return;
    }
    const int64_t pel = (int64_t)( by*8 + be.g.r ) * P.stride + bx*8;
    be.fenc = *(const uint2 *)( J.fenc + pel );
    be.ref0 = J.ref + pel; be.ref1 = NULL; be.ref_w = J.ref_w + pel;
    search_block( be, J, P, tab, bx, by, blk );
}

#define T_TILE_W 32
#define T_TILE_H 4
#ifndef X264CU_MINB_SEARCH
#define X264CU_MINB_SEARCH 5
#endif
#ifndef X264CU_MINB_FINAL
#define X264CU_MINB_FINAL 4
#endif
template<bool SATD, bool WEIGHTED, int PITCH>
__global__ void __launch_bounds__(T_TILE_W*T_TILE_H, X264CU_MINB_SEARCH) search_kernel_t( const SearchJob *__restrict__ jobs, LaParams P )
{
    __shared__ uint16_t s_tab[S_TAB_N];
    const CostTab tab = stage_cost_tab( s_tab, P );
    const SearchJob J = jobs[blockIdx.z];
    const int bx = blockIdx.x * T_TILE_W + ( threadIdx.x & ( T_TILE_W-1 ) ), by = blockIdx.y * T_TILE_H + threadIdx.x / T_TILE_W;
    if( bx >= P.mbw || by >= P.mbh ) return;
    const int blk = by * P.mbw + bx;
    if( skipped_border( P, bx, by ) ) { ( (int *)J.mv_out )[blk] = 0; J.cost_out[blk] = 0; return; }
    ThreadBE<SATD, WEIGHTED, PITCH> be; be.P = &P;
    const int pel = ( by*8 ) * be.pitch() + bx*8;
    const uint8_t *src = J.fenc + pel;
#pragma unroll
    for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * be.pitch() ) );
    be.ref0 = J.ref + pel; be.ref1 = NULL; be.ref_w = J.ref_w + pel;
    search_block( be, J, P, tab, bx, by, blk );
}

// Exact raster-serial mode: the reference evaluates blocks in reverse raster order and
// each block reads the MVs of its right / below / below-left / below-right neighbours from the SAME pass
// (slicetype.c:662-680,830-833).  Blocks with equal w = (mbw-1-x) + 2*(mbh-1-y) depend only on blocks of smaller w,
// so one launch per anti-diagonal wave reproduces the serial result exactly; mv_in == mv_out (in place).
__global__ void __launch_bounds__(CTA_THREADS) search_kernel_wave( const SearchJob *__restrict__ jobs, LaParams P, int wave )
{
    __shared__ uint16_t s_tab[S_TAB_N];
    const CostTab tab = stage_cost_tab( s_tab, P );
    const SearchJob J = jobs[blockIdx.y];
    GroupBE be; be.g.r = threadIdx.x & 7; be.g.mask = 0xffu << ( threadIdx.x & 24 ); be.P = &P;
    const int by = blockIdx.x * GROUPS_PER_CTA + ( threadIdx.x >> 3 );       // one candidate block per row
    if( by >= P.mbh ) return;
    const int bx = P.mbw - 1 - ( wave - 2 * ( P.mbh - 1 - by ) );
    if( bx < 0 || bx >= P.mbw ) return;
    if( skipped_border( P, bx, by ) ) return;                                 // the field was zeroed before the first wave
    const int blk = by * P.mbw + bx;
    const int64_t pel = (int64_t)( by*8 + be.g.r ) * P.stride + bx*8;
    be.fenc = *(const uint2 *)( J.fenc + pel );
    be.ref0 = J.ref + pel; be.ref1 = NULL; be.ref_w = J.ref_w + pel;
    search_block( be, J, P, tab, bx, by, blk );
}

void launch_search_waves( const SearchJob *d_jobs, int njobs, const LaParams &P, cudaStream_t st )
{
    dim3 grid( ( P.mbh + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, njobs );
    const int nwaves = P.mbw + 2 * ( P.mbh - 1 );
    for( int w = 0; w < nwaves; w++ )
        search_kernel_wave<<<grid, CTA_THREADS, 0, st>>>( d_jobs, P, w );
}

static bool use_thread_mapping( const LaParams &P, int njobs )
{
    if( P.lanes_per_block == 1 ) return true;
    if( P.lanes_per_block == 8 ) return false;
    return (int64_t)njobs * P.nmb >= 4 * 8160;     // wide launches: enough blocks to fill the SMs one thread each
}

template<int PITCH>
static void launch_search_t( const SearchJob *d_jobs, int njobs, bool any_weighted, const LaParams &P, cudaStream_t st )
{
    dim3 grid( ( P.mbw + T_TILE_W - 1 ) / T_TILE_W, ( P.mbh + T_TILE_H - 1 ) / T_TILE_H, njobs );
    if( P.b_satd )
    {
        if( any_weighted ) search_kernel_t<true, true, PITCH><<<grid, T_TILE_W*T_TILE_H, 0, st>>>( d_jobs, P );
        else               search_kernel_t<true, false, PITCH><<<grid, T_TILE_W*T_TILE_H, 0, st>>>( d_jobs, P );
    }
    else
    {
        if( any_weighted ) search_kernel_t<false, true, PITCH><<<grid, T_TILE_W*T_TILE_H, 0, st>>>( d_jobs, P );
        else               search_kernel_t<false, false, PITCH><<<grid, T_TILE_W*T_TILE_H, 0, st>>>( d_jobs, P );
    }
}

void launch_search( const SearchJob *d_jobs, int njobs, bool any_weighted, const LaParams &P, cudaStream_t st )
{
    if( use_thread_mapping( P, njobs ) )
    {   // the context picks the pitch from {1024, 2048} whenever the padded width fits (capi.cu), anything wider runs the
        // run-time-pitch instantiation
        if( P.stride == 1024 )      launch_search_t<1024>( d_jobs, njobs, any_weighted, P, st );
        else if( P.stride == 2048 ) launch_search_t<2048>( d_jobs, njobs, any_weighted, P, st );
        else                        launch_search_t<0>( d_jobs, njobs, any_weighted, P, st );
    }
    else
    {
        dim3 grid( ( P.nmb + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, njobs );
        search_kernel_g<<<grid, CTA_THREADS, 0, st>>>( d_jobs, P );
    }
}

// ---- mode decision + sums: the tail of slicetype_mb_cost for one block ----------------------------
template<class BE>
__device__ __forceinline__ void final_block( const BE &be, const FinalJob &J, const LaParams &P, int bx, int by, int blk )
{
    const bool border = bx == 0 || by == 0 || bx == P.mbw-1 || by == P.mbh-1;
    const bool tiny = P.mbw <= 2 || P.mbh <= 2;
    int bcost = X264CU_COST_MAX, list_used = 0;
    if( !J.b_intra_only )
    {
        int mv0x, mv0y, mv1x = 0, mv1y = 0, mc0, mc1 = X264CU_COST_MAX;
        {
            int v = __ldg( (const int *)J.mv[0] + blk );
            mv0x = (int16_t)( v & 0xffff ); mv0y = v >> 16;
            mc0 = __ldg( J.mvcost[0] + blk );
        }
        if( J.b_bidir )
        {   // slicetype.c:616-652
            int v = __ldg( (const int *)J.mv[1] + blk );
            mv1x = (int16_t)( v & 0xffff ); mv1y = v >> 16;
            mc1 = __ldg( J.mvcost[1] + blk );
            MeGeo m;
            me_limits( m, P, bx, by );
            int d0x = 0, d0y = 0, d1x = 0, d1y = 0;
            if( J.dmv_src )
            {
                int w = __ldg( (const int *)J.dmv_src + blk );
                int mrx = (int16_t)( w & 0xffff ), mry = w >> 16;
                d0x = ( mrx * J.dist_scale_factor + 128 ) >> 8;
                d0y = ( mry * J.dist_scale_factor + 128 ) >> 8;
                d1x = d0x - mrx; d1y = d0y - mry;
                d0x = iclip3( d0x, m.sminx, m.smaxx ); d0y = iclip3( d0y, m.sminy, m.smaxy );
                d1x = iclip3( d1x, m.sminx, m.smaxx ); d1y = iclip3( d1y, m.sminy, m.smaxy );
                if( !P.b_satd ) { d0x &= ~1; d0y &= ~1; d1x &= ~1; d1y &= ~1; }
            }
            // candidates in the reference's order (ties keep the earlier one): temporal direct, zero (both :644-652),
            // L0, L1 (:708), L0+L1 with penalty 5*lambda (:711-712).  The two qpel bidir costs share one call site.
#pragma unroll 1
            for( int t = 0; t < 2; t++ )
            {
                if( t == 1 )
                {
                    if( d0x | d0y | d1x | d1y )
                    {
                        int cz = be.bidir_zero( J.bipred_weight );
                        if( cz < bcost ) { bcost = cz; list_used = 3; }
                    }
                    if( mc0 < bcost ) { bcost = mc0; list_used = 1; }
                    if( mc1 < bcost ) { bcost = mc1; list_used = 2; }
                    if( !( mv0x | mv0y | mv1x | mv1y ) ) break;
                }
                int cost = ( t ? 5 : 0 ) + be.bidir( t ? mv0x : d0x, t ? mv0y : d0y, t ? mv1x : d1x, t ? mv1y : d1y, J.bipred_weight );
                if( cost < bcost ) { bcost = cost; list_used = 3; }
            }
        }
        else if( mc0 < bcost ) { bcost = mc0; list_used = 1; }
    }
    if( !be.leader() ) return;
    bcost += 4;                                         // lowres_penalty (slicetype.c:758)
    const bool score = !border || tiny;
    if( !J.b_bidir )
    {   // slicetype.c:762-773
        int icost = J.intra[blk];
        int b_intra = icost < bcost;
        if( b_intra ) { bcost = icost; list_used = 0; }
        if( score && b_intra ) atomicAdd( &J.stats[2], 1 );
    }
    if( !J.b_intra_only )
    {   // slicetype.c:776-788
        int bcost_aq = P.aq ? ( bcost * J.inv_qscale[blk] + 128 ) >> 8 : bcost;
        atomicAdd( &J.row_satds[by], bcost_aq );
        if( score )
        {
            atomicAdd( &J.stats[0], bcost );
            atomicAdd( &J.stats[1], bcost_aq );
        }
    }
    J.lowres_costs[blk] = (uint16_t)( min( bcost, 0x3FFF ) + ( list_used << 14 ) );     // slicetype.c:790
}

__global__ void __launch_bounds__(CTA_THREADS) final_kernel_g( const FinalJob *__restrict__ jobs, LaParams P )
{
    const FinalJob J = jobs[blockIdx.y];
    GroupBE be; be.g.r = threadIdx.x & 7; be.g.mask = 0xffu << ( threadIdx.x & 24 ); be.P = &P;
    const int blk = blockIdx.x * GROUPS_PER_CTA + ( threadIdx.x >> 3 );
    if( blk >= P.nmb ) return;
    const int bx = blk % P.mbw, by = blk / P.mbw;
    if( skipped_border( P, bx, by ) )
    {
        if( be.g.r == 0 && !J.b_intra_only ) J.lowres_costs[blk] = 0;
        return;
    }
    const int64_t pel = (int64_t)( by*8 + be.g.r ) * P.stride + bx*8;
    be.fenc = *(const uint2 *)( J.fenc + pel );
    be.ref0 = J.ref0 + pel; be.ref1 = J.ref1 + pel; be.ref_w = be.ref0;
    final_block( be, J, P, bx, by, blk );
}

template<bool SATD, int PITCH>
__global__ void __launch_bounds__(T_TILE_W*T_TILE_H, X264CU_MINB_FINAL) final_kernel_t( const FinalJob *__restrict__ jobs, LaParams P )
{
    const FinalJob J = jobs[blockIdx.z];
    const int bx = blockIdx.x * T_TILE_W + ( threadIdx.x & ( T_TILE_W-1 ) ), by = blockIdx.y * T_TILE_H + threadIdx.x / T_TILE_W;
    if( bx >= P.mbw || by >= P.mbh ) return;
    const int blk = by * P.mbw + bx;
    if( skipped_border( P, bx, by ) ) { if( !J.b_intra_only ) J.lowres_costs[blk] = 0; return; }
    ThreadBE<SATD, false, PITCH> be; be.P = &P;
    const int pel = ( by*8 ) * be.pitch() + bx*8;
    const uint8_t *src = J.fenc + pel;
    if( !J.b_intra_only )
    {
#pragma unroll
        for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * be.pitch() ) );
    }
    be.ref0 = J.ref0 + pel; be.ref1 = J.ref1 + pel; be.ref_w = be.ref0;
    final_block( be, J, P, bx, by, blk );
}

void launch_final( const FinalJob *d_jobs, int njobs, const LaParams &P, cudaStream_t st )
{
    if( use_thread_mapping( P, njobs ) )
    {
        dim3 grid( ( P.mbw + T_TILE_W - 1 ) / T_TILE_W, ( P.mbh + T_TILE_H - 1 ) / T_TILE_H, njobs );
#define FINAL_T( PITCH ) { if( P.b_satd ) final_kernel_t<true, PITCH><<<grid, T_TILE_W*T_TILE_H, 0, st>>>( d_jobs, P ); \
                           else           final_kernel_t<false, PITCH><<<grid, T_TILE_W*T_TILE_H, 0, st>>>( d_jobs, P ); }
        if( P.stride == 1024 )      FINAL_T( 1024 )
        else if( P.stride == 2048 ) FINAL_T( 2048 )
        else                        FINAL_T( 0 )
#undef FINAL_T
    }
    else
    {
        dim3 grid( ( P.nmb + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, njobs );
        final_kernel_g<<<grid, CTA_THREADS, 0, st>>>( d_jobs, P );
    }
}

// ------------------------------------------------------------------------------------------------
// result gather: one CTA per requested array, 128-bit copies when both ends allow it
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel( const GatherSeg *__restrict__ segs, uint8_t *__restrict__ dst_base )
{
    const GatherSeg s = segs[blockIdx.x];
    const uint8_t *src = s.src;
    uint8_t *dst = dst_base + s.dst_off;
    if( ( ( (uintptr_t)src | (uintptr_t)dst ) & 15 ) == 0 )
    {
        const unsigned n16 = s.bytes >> 4;
        for( unsigned i = threadIdx.x; i < n16; i += blockDim.x ) ( (uint4 *)dst )[i] = __ldg( (const uint4 *)src + i );
        for( unsigned i = ( n16 << 4 ) + threadIdx.x; i < s.bytes; i += blockDim.x ) dst[i] = src[i];
    }
    else if( ( ( (uintptr_t)src | (uintptr_t)dst ) & 3 ) == 0 )
    {
        const unsigned n4 = s.bytes >> 2;
        for( unsigned i = threadIdx.x; i < n4; i += blockDim.x ) ( (u32 *)dst )[i] = __ldg( (const u32 *)src + i );
        for( unsigned i = ( n4 << 2 ) + threadIdx.x; i < s.bytes; i += blockDim.x ) dst[i] = src[i];
    }
    else
        for( unsigned i = threadIdx.x; i < s.bytes; i += blockDim.x ) dst[i] = src[i];
}

void launch_gather( const GatherSeg *d_segs, int nsegs, uint8_t *dst_base, cudaStream_t st )
{
    if( nsegs > 0 ) gather_kernel<<<nsegs, 256, 0, st>>>( d_segs, dst_base );
}
