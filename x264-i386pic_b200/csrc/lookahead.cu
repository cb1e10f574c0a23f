// lookahead.cu -- sm_100a kernels of the lowres lookahead cost engine.
//
//   lowres_kernel     frame_init_lowres_core + border (mc.c:458-507, frame.c:535-555,627-631)
//   weight_kernel     x264_weight_scale_plane of the reference F plane (frame.c:825-840, mc.c:117-137)
//   intra_kernel      intra section of slicetype_mb_cost (slicetype.c:714-757)
//   gather_kernel     result arrays -> one contiguous block for a single PCIe transfer
// The motion search and decision kernels live in search.cu / final.cu, their shared block machinery in la_block.cuh.
#include "la_block.cuh"


// ------------------------------------------------------------------------------------------------
// lowres
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lum( const uint8_t *luma, int ls, int W, int H, int x, int y )
{   // mod-16 padding (frame.c:640-668) and the +1 row/column dup (mc.c:465-468) are pure replication
    return luma[(size_t)min( y, H-1 ) * ls + min( x, W-1 )];
}
#define FILT(a,b,c,d) ((((((a)+(b)+1)>>1)+(((c)+(d)+1)>>1)+1)>>1))

// The per-frame resets ride along: the evaluation statistics (x264_frame_init_lowres resets i_cost_est / i_row_satds,
// mc.c:473-481; ours are accumulated with atomics and must start at zero), the raw pixel sums of the AQ kernel and the
// default i_inv_qscale_factor of 256 (memset_int16, downscale.cl:132) -- three tiny launches in round 1.
struct FrameClear { u32 *stats; int nstats; u32 *invq2; int ninvq2; unsigned long long *pix; };

__global__ void __launch_bounds__(128) lowres_kernel( const uint8_t *__restrict__ luma, int ls, int W, int H,
                                                      uint8_t *__restrict__ planes, LaParams P, FrameClear fc )
{
    const int pw4 = ( P.lw + 2*X264CU_PAD ) >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y;
    {
        const int t = ( blockIdx.y * gridDim.x + blockIdx.x ) * blockDim.x + threadIdx.x, nt = gridDim.x * gridDim.y * blockDim.x;
        for( int i = t; i < fc.nstats; i += nt ) fc.stats[i] = 0;
        for( int i = t; i < fc.ninvq2; i += nt ) fc.invq2[i] = 0x01000100u;
        if( fc.pix && t < 8 ) fc.pix[t] = 0;
    }
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

// stats / nstats: 32-bit words to zero; inv_qscale != NULL: fill nmb entries (rounded up to a pair; the array has slack) with 256;
// pixstats != NULL: zero its 8 words
void launch_lowres( const uint8_t *luma, int stride, int width, int height, uint8_t *planes, const LaParams &P,
                    int32_t *stats, int nstats, uint16_t *inv_qscale, unsigned long long *pixstats, cudaStream_t st )
{
    int pw4 = ( P.lw + 2*X264CU_PAD ) >> 2;
    dim3 grid( ( pw4 + 127 ) / 128, P.lh + 2*X264CU_PAD );
    FrameClear fc = { (u32 *)stats, stats ? nstats : 0, (u32 *)inv_qscale, inv_qscale ? ( P.nmb + 1 ) / 2 : 0, pixstats };
    lowres_kernel<<<grid, 128, 0, st>>>( luma, stride, width, height, planes, P, fc );
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
