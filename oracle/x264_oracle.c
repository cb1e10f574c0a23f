/* oracle/x264_oracle.c -- TEST INFRASTRUCTURE ONLY (see x264_oracle.h).
 *
 * Plain-C restatement of the x264 lookahead cost path and the batched pixel primitives, written
 * from the behaviour of the reference (x264 build 164 as vendored in xrgtn/x264-i386pic); each
 * function cites the reference file:line it follows.  It is deliberately scalar and obvious:
 * plain ints instead of the reference's packed 16-bit lanes (the lanes cannot overflow for 8-bit
 * input; tests/test_oracle_vs_ref.py checks that claim on checkasm's overflow vectors).
 *
 * The lookahead uses the DEPENDENCY-FREE schedule that the GPU implements: the per-block
 * computation is the reference's slicetype_mb_cost (slicetype.c:514-791), but the neighbour MVs
 * that seed the predictor (slicetype.c:662-680) are taken from the previous full-frame pass
 * (pass 0: the zero field) instead of from blocks finished earlier in the same raster pass.
 * After `search_passes` passes the last pass's MVs/costs are the result.
 */
#include "x264_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define XO_PAD 32              /* PADH == PADV == 32, frame.h:32-33 */
#define XO_COST_MAX (1<<28)    /* COST_MAX, encoder/me.h */
#define XO_MAXB 16             /* X264_BFRAME_MAX */
#define XO_MAX_FRAMES 512

static inline int imin( int a, int b ) { return a < b ? a : b; }
static inline int imax( int a, int b ) { return a > b ? a : b; }
static inline int clip3( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }
static inline int clip_pixel( int v ) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int median3( int a, int b, int c )
{   /* x264_median, common/base.h */
    int t = (a-b)&((a-b)>>31);
    a -= t; b += t;
    b -= (b-c)&((b-c)>>31);
    b += (a-b)&((a-b)>>31);
    return b;
}

static const int blk_w[8] = { 16, 16, 8, 8, 8, 4, 4, 4 };
static const int blk_h[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

/* ================================================================================================ */
/* primitives                                                                                       */
/* ================================================================================================ */

/* pixel.c:55-80 */
int xo_sad( int size, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int w = blk_w[size], h = blk_h[size], s = 0;
    for( int y = 0; y < h; y++, a += sa, b += sb )
        for( int x = 0; x < w; x++ )
            s += abs( a[x] - b[x] );
    return s;
}

/* unnormalised 4x4 Hadamard of a difference block: rows then columns (pixel.c:242-251,265-288) */
static void hadamard4x4_diff( int c[16], const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int t[16];
    for( int y = 0; y < 4; y++, a += sa, b += sb )
    {
        int d0 = a[0]-b[0], d1 = a[1]-b[1], d2 = a[2]-b[2], d3 = a[3]-b[3];
        int s01 = d0+d1, m01 = d0-d1, s23 = d2+d3, m23 = d2-d3;
        t[y*4+0] = s01+s23; t[y*4+1] = m01+m23; t[y*4+2] = s01-s23; t[y*4+3] = m01-m23;
    }
    for( int x = 0; x < 4; x++ )
    {
        int s01 = t[x]+t[4+x], m01 = t[x]-t[4+x], s23 = t[8+x]+t[12+x], m23 = t[8+x]-t[12+x];
        c[x] = s01+s23; c[4+x] = m01+m23; c[8+x] = s01-s23; c[12+x] = m01-m23;
    }
}

static int sum_abs16( const int c[16] )
{
    int s = 0;
    for( int i = 0; i < 16; i++ ) s += abs( c[i] );
    return s;
}

/* pixel.c:265 (4x4: sum>>1) and :290 (8x4: both halves summed, then >>1) */
static int satd_4x4( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int c[16];
    hadamard4x4_diff( c, a, sa, b, sb );
    return sum_abs16( c ) >> 1;
}
static int satd_8x4( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int c[16], s;
    hadamard4x4_diff( c, a, sa, b, sb );     s  = sum_abs16( c );
    hadamard4x4_diff( c, a+4, sa, b+4, sb ); s += sum_abs16( c );
    return s >> 1;
}

/* pixel.c:311-332: widths >= 8 are tiled with 8x4, width 4 with 4x4 */
int xo_satd( int size, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int w = blk_w[size], h = blk_h[size], s = 0;
    for( int y = 0; y < h; y += 4 )
        for( int x = 0; x < w; x += ( w == 4 ? 4 : 8 ) )
            s += w == 4 ? satd_4x4( a+x+y*sa, sa, b+x+y*sb, sb ) : satd_8x4( a+x+y*sa, sa, b+x+y*sb, sb );
    return s;
}

/* 8-point Hadamard butterflies in place, stride st */
static void hadamard8( int *v, int st )
{
    for( int span = 1; span < 8; span <<= 1 )
        for( int i = 0; i < 8; i++ )
            if( !(i & span) )
            {
                int p = v[i*st], q = v[(i+span)*st];
                v[i*st] = p+q; v[(i+span)*st] = p-q;
            }
}

/* pixel.c:334-366: un-normalised sum of |8x8 Hadamard| */
static int sa8d_8x8_raw( const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int d[64], s = 0;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
            d[y*8+x] = a[x+y*sa] - b[x+y*sb];
    for( int y = 0; y < 8; y++ ) hadamard8( d+y*8, 1 );
    for( int x = 0; x < 8; x++ ) hadamard8( d+x, 8 );
    for( int i = 0; i < 64; i++ ) s += abs( d[i] );
    return s;
}

/* pixel.c:368-381 */
int xo_sa8d( int size, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    if( size == XO_8x8 )
        return ( sa8d_8x8_raw( a, sa, b, sb ) + 2 ) >> 2;
    if( size == XO_16x16 )
        return ( sa8d_8x8_raw( a, sa, b, sb ) + sa8d_8x8_raw( a+8, sa, b+8, sb )
               + sa8d_8x8_raw( a+8*sa, sa, b+8*sb, sb ) + sa8d_8x8_raw( a+8+8*sa, sa, b+8+8*sb, sb ) + 2 ) >> 2;
    return -1;
}

/* pixel.c:383-418: per 8x8 of the SOURCE: sum|4x4 Hadamard| and sum|8x8 Hadamard|, each minus DC */
static void hadamard_ac_8x8( const uint8_t *p, intptr_t st, uint32_t *sum4, uint32_t *sum8 )
{
    static const uint8_t zero[4] = { 0, 0, 0, 0 };
    int c[4][16], s4 = 0, s8 = 0, dc = 0;
    for( int k = 0; k < 4; k++ )
    {
        hadamard4x4_diff( c[k], p + (k&1)*4 + (k>>1)*4*st, st, zero, 0 );
        s4 += sum_abs16( c[k] );
        dc += c[k][0];
    }
    for( int i = 0; i < 16; i++ )
    {
        int a0 = c[0][i]+c[1][i], a1 = c[0][i]-c[1][i], a2 = c[2][i]+c[3][i], a3 = c[2][i]-c[3][i];
        s8 += abs( a0+a2 ) + abs( a0-a2 ) + abs( a1+a3 ) + abs( a1-a3 );
    }
    *sum4 += s4 - dc;
    *sum8 += s8 - dc;
}

/* pixel.c:420-435 */
uint64_t xo_hadamard_ac( int size, const uint8_t *p, intptr_t st )
{
    if( size > XO_8x8 ) return ~0ULL;
    uint32_t s4 = 0, s8 = 0;
    for( int y = 0; y < blk_h[size]; y += 8 )
        for( int x = 0; x < blk_w[size]; x += 8 )
            hadamard_ac_8x8( p + x + y*st, st, &s4, &s8 );
    return ( (uint64_t)( s8 >> 2 ) << 32 ) + ( s4 >> 1 );
}

int xo_pixel_cmp_batch( int kind, int size, const uint8_t *base1, intptr_t s1, const uint8_t *base2, intptr_t s2,
                        int64_t n, const int64_t *off1, const int64_t *off2, int32_t *out )
{
    for( int64_t i = 0; i < n; i++ )
    {
        const uint8_t *a = base1 + off1[i], *b = base2 + off2[i];
        out[i] = kind == 0 ? xo_sad( size, a, s1, b, s2 ) : kind == 1 ? xo_satd( size, a, s1, b, s2 )
               : kind == 2 ? xo_sa8d( size, a, s1, b, s2 ) : -1;
    }
    return 0;
}

int xo_hadamard_ac_batch( int size, const uint8_t *base, intptr_t stride, int64_t n, const int64_t *off, uint64_t *out )
{
    for( int64_t i = 0; i < n; i++ )
        out[i] = xo_hadamard_ac( size, base + off[i], stride );
    return 0;
}

/* dct.c:157-189: H.264 core transform of the residual, rows then columns, no scaling.
 * Output order follows the reference: first index = vertical frequency?  No: dct[i*4+j] with the
 * second pass writing dct[i*4+0..3] from tmp[i*4+..], where tmp[k*4+i] was written by pass one --
 * i.e. the block comes out transposed exactly as in the reference. */
static void sub4x4_dct( int16_t dct[16], const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    int d[16], tmp[16];
    for( int y = 0; y < 4; y++ )
        for( int x = 0; x < 4; x++ )
            d[y*4+x] = a[x+y*sa] - b[x+y*sb];
    for( int i = 0; i < 4; i++ )
    {
        int s03 = d[i*4+0]+d[i*4+3], s12 = d[i*4+1]+d[i*4+2];
        int d03 = d[i*4+0]-d[i*4+3], d12 = d[i*4+1]-d[i*4+2];
        tmp[0*4+i] = (int16_t)(   s03 +   s12 );
        tmp[1*4+i] = (int16_t)( 2*d03 +   d12 );
        tmp[2*4+i] = (int16_t)(   s03 -   s12 );
        tmp[3*4+i] = (int16_t)(   d03 - 2*d12 );
    }
    for( int i = 0; i < 4; i++ )
    {
        int s03 = tmp[i*4+0]+tmp[i*4+3], s12 = tmp[i*4+1]+tmp[i*4+2];
        int d03 = tmp[i*4+0]-tmp[i*4+3], d12 = tmp[i*4+1]-tmp[i*4+2];
        dct[i*4+0] = (int16_t)(   s03 +   s12 );
        dct[i*4+1] = (int16_t)( 2*d03 +   d12 );
        dct[i*4+2] = (int16_t)(   s03 -   s12 );
        dct[i*4+3] = (int16_t)(   d03 - 2*d12 );
    }
}

/* dct.c:191-205: 8x8 = four 4x4 in Z order; 16x16 = four 8x8 in Z order */
void xo_sub_dct( int bsize, int16_t *out, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb )
{
    if( bsize == 4 )
        sub4x4_dct( out, a, sa, b, sb );
    else
        for( int k = 0; k < 4; k++ )
        {
            int h = bsize/2, ox = (k&1)*h, oy = (k>>1)*h;
            xo_sub_dct( h, out + k*h*h, a + ox + oy*sa, sa, b + ox + oy*sb, sb );
        }
}

int xo_sub_dct_batch( int bsize, const uint8_t *base1, intptr_t s1, const uint8_t *base2, intptr_t s2,
                      int64_t n, const int64_t *off1, const int64_t *off2, int16_t *out )
{
    if( bsize != 4 && bsize != 8 && bsize != 16 ) return -1;
    for( int64_t i = 0; i < n; i++ )
        xo_sub_dct( bsize, out + i*bsize*bsize, base1 + off1[i], s1, base2 + off2[i], s2 );
    return 0;
}

/* quant.c:50-73 */
int xo_quant( int16_t *coef, int n, const uint16_t *mf, const uint16_t *bias )
{
    int nz = 0;
    for( int i = 0; i < n; i++ )
    {
        int c = coef[i];
        if( c > 0 ) c = (int)( ( (uint32_t)bias[i] + (uint32_t)c ) * mf[i] >> 16 );
        else        c = -(int32_t)( ( (uint32_t)bias[i] + (uint32_t)(-c) ) * mf[i] >> 16 );
        coef[i] = (int16_t)c;
        nz |= (int16_t)c;
    }
    return !!nz;
}

int xo_quant_batch( int b8x8, int16_t *coefs, int64_t nblocks, const uint16_t *mf, const uint16_t *bias, int32_t *nz )
{
    int n = b8x8 ? 64 : 16;
    for( int64_t i = 0; i < nblocks; i++ )
        nz[i] = xo_quant( coefs + i*n, n, mf, bias );
    return 0;
}

/* mc.c:484-507 */
void xo_lowres_core( const uint8_t *src, intptr_t ss, uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc,
                     intptr_t ds, int width, int height )
{
#define AVG2(a,b) (((a)+(b)+1)>>1)
#define FILT(a,b,c,d) AVG2( AVG2(a,b), AVG2(c,d) )
    for( int y = 0; y < height; y++ )
    {
        const uint8_t *r0 = src + 2*y*ss, *r1 = r0 + ss, *r2 = r1 + ss;
        for( int x = 0; x < width; x++ )
        {
            d0[x+y*ds] = FILT( r0[2*x  ], r1[2*x  ], r0[2*x+1], r1[2*x+1] );
            dh[x+y*ds] = FILT( r0[2*x+1], r1[2*x+1], r0[2*x+2], r1[2*x+2] );
            dv[x+y*ds] = FILT( r1[2*x  ], r2[2*x  ], r1[2*x+1], r2[2*x+1] );
            dc[x+y*ds] = FILT( r1[2*x+1], r2[2*x+1], r1[2*x+2], r2[2*x+2] );
        }
    }
}

/* tables.c:183-184 */
static const uint8_t hpel_ref0[16] = {0,1,1,1,0,1,1,1,2,3,3,3,0,1,1,1};
static const uint8_t hpel_ref1[16] = {0,0,1,0,2,2,3,2,2,2,3,2,2,2,3,2};

/* mc.c:117-137 */
static inline int weight_px( int v, int scale, int denom, int offset )
{
    if( denom >= 1 ) return clip_pixel( ( ( v*scale + (1<<(denom-1)) ) >> denom ) + offset );
    return clip_pixel( v*scale + offset );
}

/* mc.c:198-249 (mc_luma == get_ref made to always write dst); denom < 0: unweighted */
void xo_mc_luma( uint8_t *dst, intptr_t ds, const uint8_t *const planes[4], intptr_t stride,
                 int mvx, int mvy, int w, int h, int scale, int denom, int offset )
{
    int qidx = ((mvy&3)<<2) + (mvx&3);
    intptr_t off = (intptr_t)(mvy>>2)*stride + (mvx>>2);
    const uint8_t *s1 = planes[hpel_ref0[qidx]] + off + ((mvy&3) == 3)*stride;
    const uint8_t *s2 = planes[hpel_ref1[qidx]] + off + ((mvx&3) == 3);
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
        {
            int v = s1[x+y*stride];
            if( qidx & 5 ) v = ( v + s2[x+y*stride] + 1 ) >> 1;
            if( denom >= 0 ) v = weight_px( v, scale, denom, offset );
            dst[x+y*ds] = v;
        }
}

/* mc.c:63-99 */
void xo_avg( uint8_t *dst, intptr_t ds, const uint8_t *a, intptr_t as, const uint8_t *b, intptr_t bs,
             int w, int h, int weight )
{
    for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
            dst[x+y*ds] = weight == 32 ? ( a[x+y*as] + b[x+y*bs] + 1 ) >> 1
                        : clip_pixel( ( a[x+y*as]*weight + b[x+y*bs]*(64-weight) + 32 ) >> 6 );
}

/* mc.c:119-137 mc_weight over a width x height area */
void xo_mc_weight( uint8_t *dst, intptr_t ds, const uint8_t *src, intptr_t ss, int width, int height, int scale, int denom, int offset )
{
    for( int y = 0; y < height; y++ )
        for( int x = 0; x < width; x++ )
            dst[x+y*ds] = weight_px( src[x+y*ss], scale, denom, offset );
}

/* analyse.c:153-157,186-188 with lambda = x264_lambda_tab[X264_LOOKAHEAD_QP] = 1 (tables.c:97) */
void xo_cost_mv_table( uint16_t *centre, int n )
{
    for( int i = 0; i <= n; i++ )
    {
        float l = i ? log2f( i+1 ) * 2.0f + 1.718f : 0.718f;
        int c = imin( (int)( 1 * l + .5f ), 65535 );
        centre[i] = centre[-i] = c;
    }
}

/* ---- intra: predict.c:221-304 (8x8c DC/H/V/P), :632-676 (edge filter), :741-887 (directional) ---- */
#define F1(a,b)   (((a)+(b)+1)>>1)
#define F2(a,b,c) (((a)+2*(b)+(c)+2)>>2)

/* top[-1..15], left[-1..7] (left[-1] == top[-1] == top-left) */
static void pred8x8c_dc( uint8_t p[64], const uint8_t *top, const uint8_t *left )
{
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for( int i = 0; i < 4; i++ ) { s0 += top[i]; s1 += top[i+4]; s2 += left[i]; s3 += left[i+4]; }
    int dc[4] = { (s0+s2+4)>>3, (s1+2)>>2, (s3+2)>>2, (s1+s3+4)>>3 };
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
            p[y*8+x] = dc[(y>>2)*2+(x>>2)];
}
static void pred8x8c_h( uint8_t p[64], const uint8_t *left )
{
    for( int y = 0; y < 8; y++ ) memset( p+y*8, left[y], 8 );
}
static void pred8x8c_v( uint8_t p[64], const uint8_t *top )
{
    for( int y = 0; y < 8; y++ ) memcpy( p+y*8, top, 8 );
}
static void pred8x8c_p( uint8_t p[64], const uint8_t *top, const uint8_t *left )
{
    int H = 0, V = 0;
    for( int i = 0; i < 4; i++ )
    {
        H += (i+1) * ( top[4+i] - top[2-i] );       /* top[-1] for i == 3 */
        V += (i+1) * ( left[4+i] - left[2-i] );
    }
    int a = 16 * ( left[7] + top[7] );
    int b = ( 17*H + 16 ) >> 5;
    int c = ( 17*V + 16 ) >> 5;
    int i00 = a - 3*b - 3*c + 16;
    for( int y = 0; y < 8; y++, i00 += c )
        for( int x = 0, pix = i00; x < 8; x++, pix += b )
            p[y*8+x] = clip_pixel( pix >> 5 );
}

typedef struct { int lt, l[8], t[16]; } edge_t;

/* predict.c:632-676 with i_neighbor = i_filters = ALL_NEIGHBORS */
static void filter_edge( edge_t *e, const uint8_t *top, const uint8_t *left )
{
    int tl = top[-1];
    e->lt = ( top[0] + 2*tl + left[0] + 2 ) >> 2;
    e->l[0] = ( tl + 2*left[0] + left[1] + 2 ) >> 2;
    for( int y = 1; y < 7; y++ ) e->l[y] = F2( left[y-1], left[y], left[y+1] );
    e->l[7] = ( left[6] + 3*left[7] + 2 ) >> 2;
    e->t[0] = ( tl + 2*top[0] + top[1] + 2 ) >> 2;
    for( int x = 1; x < 15; x++ ) e->t[x] = F2( top[x-1], top[x], top[x+1] );
    e->t[15] = ( top[14] + 3*top[15] + 2 ) >> 2;
}

#define P(x,y) p[(y)*8+(x)]
static void pred8x8_ddl( uint8_t p[64], const edge_t *e )
{
    const int *t = e->t;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int i = x+y;
            P(x,y) = i < 14 ? F2( t[i], t[i+1], t[i+2] ) : F2( t[14], t[15], t[15] );
        }
}
static void pred8x8_ddr( uint8_t p[64], const edge_t *e )
{
    /* one diagonal array: v[0..7] = l7..l0, v[8] = lt, v[9..16] = t0..t7 */
    int v[17];
    for( int i = 0; i < 8; i++ ) { v[i] = e->l[7-i]; v[9+i] = e->t[i]; }
    v[8] = e->lt;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int k = 8 + x - y;          /* centre tap */
            P(x,y) = F2( v[k-1], v[k], v[k+1] );
        }
}
static void pred8x8_vr( uint8_t p[64], const edge_t *e )
{
    int v[17];
    for( int i = 0; i < 8; i++ ) { v[i] = e->l[7-i]; v[9+i] = e->t[i]; }
    v[8] = e->lt;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int z = 2*x - y;
            if( z >= 0 )
            {
                int k = 8 + x - (y>>1);           /* v[k] = t[x-(y>>1)-1] or lt */
                P(x,y) = (z&1) ? F2( v[k-1], v[k], v[k+1] ) : F1( v[k], v[k+1] );
            }
            else if( z == -1 )
                P(x,y) = F2( v[7], v[8], v[9] );  /* F2(l0,lt,t0) */
            else
            {
                int k = 8 - (y - 2*x - 1);        /* centre = l[y-2x-2] */
                P(x,y) = F2( v[k-1], v[k], v[k+1] );
            }
        }
}
static void pred8x8_hd( uint8_t p[64], const edge_t *e )
{
    int v[17];
    for( int i = 0; i < 8; i++ ) { v[i] = e->l[7-i]; v[9+i] = e->t[i]; }
    v[8] = e->lt;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int z = 2*y - x;
            if( z >= 0 )
            {
                int k = 8 - (y - (x>>1));         /* v[k] = l[y-(x>>1)-1] or lt */
                P(x,y) = (z&1) ? F2( v[k-1], v[k], v[k+1] ) : F1( v[k-1], v[k] );
            }
            else if( z == -1 )
                P(x,y) = F2( v[7], v[8], v[9] );
            else
            {
                int k = 8 + (x - 2*y - 1);        /* centre = t[x-2y-2] */
                P(x,y) = F2( v[k-1], v[k], v[k+1] );
            }
        }
}
static void pred8x8_vl( uint8_t p[64], const edge_t *e )
{
    const int *t = e->t;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int i = x + (y>>1);
            P(x,y) = (y&1) ? F2( t[i], t[i+1], t[i+2] ) : F1( t[i], t[i+1] );
        }
}
static void pred8x8_hu( uint8_t p[64], const edge_t *e )
{
    const int *l = e->l;
    for( int y = 0; y < 8; y++ )
        for( int x = 0; x < 8; x++ )
        {
            int z = x + 2*y;
            if( z > 13 )       P(x,y) = l[7];
            else if( z == 13 ) P(x,y) = F2( l[6], l[7], l[7] );
            else if( z & 1 )   P(x,y) = F2( l[z>>1], l[(z>>1)+1], l[(z>>1)+2] );
            else               P(x,y) = F1( l[z>>1], l[(z>>1)+1] );
        }
}
#undef P

static int cmp8x8( int b_satd, const uint8_t *fenc, intptr_t sf, const uint8_t *pred )
{
    return b_satd ? xo_satd( XO_8x8, fenc, sf, pred, 8 ) : xo_sad( XO_8x8, fenc, sf, pred, 8 );
}

/* slicetype.c:714-746 up to (not including) the +5*lambda+4: min over modes of mbcmp 8x8.
 * src points at the block inside a plane that has valid pixels around it (32-px border). */
static int intra_min_cost( const uint8_t *src, intptr_t st, int b_satd )
{
    uint8_t topbuf[17], leftbuf[9], p[64];
    uint8_t *top = topbuf+1, *left = leftbuf+1;
    for( int i = -1; i < 16; i++ ) top[i] = src[i - st];
    for( int i = -1; i < 8; i++ )  left[i] = src[-1 + i*st];
    int best;
    pred8x8c_dc( p, top, left ); best = cmp8x8( b_satd, src, st, p );
    pred8x8c_h( p, left );       best = imin( best, cmp8x8( b_satd, src, st, p ) );
    pred8x8c_v( p, top );        best = imin( best, cmp8x8( b_satd, src, st, p ) );
    if( b_satd )   /* param subme > 1, slicetype.c:732 */
    {
        edge_t e;
        pred8x8c_p( p, top, left ); best = imin( best, cmp8x8( 1, src, st, p ) );
        filter_edge( &e, top, left );
        pred8x8_ddl( p, &e ); best = imin( best, cmp8x8( 1, src, st, p ) );
        pred8x8_ddr( p, &e ); best = imin( best, cmp8x8( 1, src, st, p ) );
        pred8x8_vr( p, &e );  best = imin( best, cmp8x8( 1, src, st, p ) );
        pred8x8_hd( p, &e );  best = imin( best, cmp8x8( 1, src, st, p ) );
        pred8x8_vl( p, &e );  best = imin( best, cmp8x8( 1, src, st, p ) );
        pred8x8_hu( p, &e );  best = imin( best, cmp8x8( 1, src, st, p ) );
    }
    return best;
}

/* intra_sad_x3_8x8c / intra_satd_x3_8x8c, pixel.c:542-556: res = { DC, H, V }; the predictors read the row above and the
 * column left of fdec (predict.c:221-277) */
void xo_intra_x3_8x8c( int b_satd, const uint8_t *fenc, intptr_t sf, const uint8_t *fdec, intptr_t sd, int res[3] )
{
    uint8_t top[17] = {0}, left[9], p[64];
    for( int i = 0; i < 8; i++ ) top[1+i] = fdec[i - sd];
    top[0] = left[0] = fdec[-1 - sd];
    for( int i = 0; i < 8; i++ ) left[1+i] = fdec[i*sd - 1];
    pred8x8c_dc( p, top + 1, left + 1 ); res[0] = cmp8x8( b_satd, fenc, sf, p );
    pred8x8c_h( p, left + 1 );           res[1] = cmp8x8( b_satd, fenc, sf, p );
    pred8x8c_v( p, top + 1 );            res[2] = cmp8x8( b_satd, fenc, sf, p );
}
int xo_intra_x3_8x8c_batch( int b_satd, const uint8_t *p1, intptr_t s1, const uint8_t *p2, intptr_t s2, int64_t n,
                            const int64_t *off1, const int64_t *off2, int32_t *res )
{
    for( int64_t i = 0; i < n; i++ )
        xo_intra_x3_8x8c( b_satd, p1 + off1[i], s1, p2 + off2[i], s2, res + 3*i );
    return 0;
}

int xo_intra_cost_8x8( const uint8_t *src, intptr_t stride, int b_satd )
{
    return intra_min_cost( src, stride, b_satd ) + 5*1 + 4;   /* slicetype.c:720,746: +5*lambda +lowres_penalty */
}

/* ================================================================================================ */
/* lookahead                                                                                        */
/* ================================================================================================ */
typedef struct
{
    int valid;
    uint8_t *buf;                 /* 4 padded planes */
    uint8_t *plane[4];            /* -> pixel (0,0) of each plane */
    uint8_t *wbuf, *wplane;       /* weighted copy of plane 0 (slicetype.c:490-500) */
    uint16_t *inv_qscale;
    int b_intra_calculated;
    uint16_t *lowres_costs[XO_MAXB+2][XO_MAXB+2];
    int16_t (*lowres_mvs[2][XO_MAXB+1])[2];
    int *lowres_mv_costs[2][XO_MAXB+1];
    int  searched[2][XO_MAXB+1];  /* stands for the 0x7FFF sentinel (mc.c:479-481, slicetype.c:855-867) */
    int *row_satds[XO_MAXB+2][XO_MAXB+2];
    int cost_est[XO_MAXB+2][XO_MAXB+2];
    int cost_est_aq[XO_MAXB+2][XO_MAXB+2];
    int intra_mbs[XO_MAXB+2];
    uint16_t *propagate;          /* i_propagate_cost */
    uint64_t pixel_sum[3], pixel_ssd[3];   /* i_pixel_sum / i_pixel_ssd after the mean removal (ratecontrol.c:406-414) */
    float *qp_offset_aq;          /* f_qp_offset_aq == f_qp_offset right after x264_adaptive_quant_frame */
    int have_aq;
} xo_frame;

struct xo_ctx
{
    xo_cfg cfg;
    int mbw, mbh, nmb, lw, lh, stride, plane_rows;
    int b_satd;                   /* param subme > 1: mbcmp = SATD (encoder.c:1409-1427) */
    int me_method, la_subme;      /* lowres_context_init, slicetype.c:45-61 */
    uint16_t *cost_mv_buf, *cost_mv;
    uint8_t *fullres;             /* scratch (W16+1) x (H16+1) */
    xo_frame fr[XO_MAX_FRAMES];
    int16_t (*snap[2])[2];
    int64_t cnt[3];
};

xo_ctx *xo_open( const xo_cfg *cfg )
{
    xo_ctx *c = calloc( 1, sizeof(*c) );
    if( !c ) return NULL;
    c->cfg = *cfg;
    if( c->cfg.search_passes < 0 ) c->cfg.search_passes = 1;      /* 0: the reference's own raster-serial schedule */
    c->mbw = ( cfg->width + 15 ) / 16;
    c->mbh = ( cfg->height + 15 ) / 16;
    c->nmb = c->mbw * c->mbh;
    c->lw = c->mbw * 8;
    c->lh = c->mbh * 8;
    c->stride = c->lw + 2*XO_PAD;
    c->plane_rows = c->lh + 2*XO_PAD;
    c->b_satd = cfg->subme > 1;
    c->me_method = c->b_satd ? imin( 1, cfg->me_method ) : 0;
    c->la_subme = c->b_satd ? 4 : 2;
    int n = 2*4*cfg->mv_range;
    c->cost_mv_buf = malloc( (2*n+1) * sizeof(uint16_t) );
    c->cost_mv = c->cost_mv_buf + n;
    xo_cost_mv_table( c->cost_mv, n );
    c->fullres = malloc( (size_t)(c->mbw*16+1) * (c->mbh*16+1) );
    for( int l = 0; l < 2; l++ ) c->snap[l] = calloc( c->nmb, 4 );
    return c;
}

static void free_frame( xo_frame *f )
{
    free( f->buf ); free( f->wbuf ); free( f->inv_qscale ); free( f->propagate ); free( f->qp_offset_aq );
    for( int i = 0; i < XO_MAXB+2; i++ )
        for( int j = 0; j < XO_MAXB+2; j++ ) { free( f->lowres_costs[i][j] ); free( f->row_satds[i][j] ); }
    for( int l = 0; l < 2; l++ )
        for( int i = 0; i < XO_MAXB+1; i++ ) { free( f->lowres_mvs[l][i] ); free( f->lowres_mv_costs[l][i] ); }
    memset( f, 0, sizeof(*f) );
}

void xo_close( xo_ctx *c )
{
    if( !c ) return;
    for( int i = 0; i < XO_MAX_FRAMES; i++ ) if( c->fr[i].valid ) free_frame( &c->fr[i] );
    free( c->cost_mv_buf ); free( c->fullres );
    for( int l = 0; l < 2; l++ ) free( c->snap[l] );
    free( c );
}

int xo_info( xo_ctx *c, int *out )
{
    out[0] = c->mbw; out[1] = c->mbh; out[2] = c->lw; out[3] = c->lh;
    return 0;
}

/* frame.c:535-555,627-631: replicate 32 px on every side */
static void expand_border( uint8_t *p, int stride, int w, int h )
{
    for( int y = 0; y < h; y++ )
    {
        memset( p + y*stride - XO_PAD, p[y*stride], XO_PAD );
        memset( p + y*stride + w, p[y*stride + w-1], XO_PAD );
    }
    for( int y = 1; y <= XO_PAD; y++ )
    {
        memcpy( p - y*stride - XO_PAD, p - XO_PAD, w + 2*XO_PAD );
        memcpy( p + (h-1+y)*stride - XO_PAD, p + (h-1)*stride - XO_PAD, w + 2*XO_PAD );
    }
}

/* encoder.c:3372-3377 (mod16 pad, frame.c:640-668) + mc.c:458-482 */
int xo_put_frame( xo_ctx *c, int idx, const uint8_t *luma, int stride, const uint16_t *inv_qscale )
{
    if( idx < 0 || idx >= XO_MAX_FRAMES ) return -1;
    xo_frame *f = &c->fr[idx];
    if( f->valid ) free_frame( f );
    int W = c->cfg.width, H = c->cfg.height, W16 = c->mbw*16, H16 = c->mbh*16, fs = W16+1;
    for( int y = 0; y <= H16; y++ )
    {
        const uint8_t *src = luma + (intptr_t)imin( y, H-1 ) * stride;
        uint8_t *dst = c->fullres + (size_t)y*fs;
        memcpy( dst, src, W );
        memset( dst + W, src[W-1], fs - W );
    }
    size_t psz = (size_t)c->stride * c->plane_rows;
    f->buf = malloc( 4*psz );
    for( int i = 0; i < 4; i++ ) f->plane[i] = f->buf + i*psz + XO_PAD*c->stride + XO_PAD;
    xo_lowres_core( c->fullres, fs, f->plane[0], f->plane[1], f->plane[2], f->plane[3], c->stride, c->lw, c->lh );
    for( int i = 0; i < 4; i++ ) expand_border( f->plane[i], c->stride, c->lw, c->lh );
    f->inv_qscale = malloc( c->nmb * sizeof(uint16_t) );
    for( int i = 0; i < c->nmb; i++ ) f->inv_qscale[i] = inv_qscale ? inv_qscale[i] : 256;
    for( int i = 0; i < XO_MAXB+2; i++ )
        for( int j = 0; j < XO_MAXB+2; j++ ) f->cost_est[i][j] = f->cost_est_aq[i][j] = -1;
    f->valid = 1;
    return 0;
}

int xo_get_lowres( xo_ctx *c, int idx, int p, uint8_t *out )
{
    xo_frame *f = &c->fr[idx];
    if( !f->valid ) return -1;
    memcpy( out, f->plane[p] - XO_PAD*c->stride - XO_PAD, (size_t)c->stride * c->plane_rows );
    return 0;
}

/* ---- motion search (encoder/me.c) ---- */
typedef struct
{
    const uint8_t *fenc; int fs;        /* 8x8 source block */
    const uint8_t *ref[4];              /* the four lowres phase planes at the block position */
    const uint8_t *ref_w;               /* fullpel plane (weighted copy if weights are on) */
    int stride;
    int wscale, wdenom, woffset;        /* wdenom < 0: none */
    int mvp[2];
    int fmin[2], fmax[2];               /* mv_limit_fpel */
    int smin[2], smax[2];               /* mv_min_spel / mv_max_spel */
    int mv[2], cost;
} me_t;

static inline int mvcost( const xo_ctx *c, const me_t *m, int mx, int my )
{   /* p_cost_mvx[mx] + p_cost_mvy[my], me.c:211-212 */
    return c->cost_mv[mx - m->mvp[0]] + c->cost_mv[my - m->mvp[1]];
}

static int sad_fpel( xo_ctx *c, const me_t *m, int mx, int my )
{   /* COST_MV without the mv bits, me.c:63-70 */
    c->cnt[0]++;
    return xo_sad( XO_8x8, m->fenc, m->fs, m->ref_w + my*m->stride + mx, m->stride );
}

static int cmp_qpel( xo_ctx *c, const me_t *m, int mx, int my, int b_satd )
{   /* COST_MV_HPEL / COST_MV_SAD / COST_MV_SATD without the mv bits, me.c:72-79,817-832 */
    uint8_t pix[64];
    xo_mc_luma( pix, 8, m->ref, m->stride, mx, my, 8, 8, m->wscale, m->wdenom, m->woffset );
    c->cnt[b_satd ? 2 : 1]++;
    return b_satd ? xo_satd( XO_8x8, m->fenc, m->fs, pix, 8 ) : xo_sad( XO_8x8, m->fenc, m->fs, pix, 8 );
}

static inline int in_range( const me_t *m, int mx, int my )
{   /* CHECK_MVRANGE, me.c:209 */
    return mx >= m->fmin[0] && mx <= m->fmax[0] && my >= m->fmin[1] && my <= m->fmax[1];
}

#define FPEL(v) (((v)+2)>>2)

static const int8_t hex2[8][2] = {{-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0}};   /* me.c:55 */
static const uint8_t mod6m1[8] = {5,0,1,2,3,4,5,0};                                                 /* me.c:53 */
static const int8_t square1[9][2] = {{0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1}}; /* me.c:56 */

/* me.c:865-992 with p_halfpel_thresh = NULL, b_refine_qpel = 0, no chroma ME */
static void refine_subpel( xo_ctx *c, me_t *m, int hpel_iters, int qpel_iters )
{
    int bmx = m->mv[0], bmy = m->mv[1], bcost = m->cost;

    if( hpel_iters )
    {
        if( c->la_subme < 3 )
        {   /* me.c:889-895 */
            int mx = clip3( m->mvp[0], m->smin[0]+2, m->smax[0]-2 );
            int my = clip3( m->mvp[1], m->smin[1]+2, m->smax[1]-2 );
            if( (mx-bmx) | (my-bmy) )
            {
                int cost = cmp_qpel( c, m, mx, my, 0 ) + mvcost( c, m, mx, my );
                if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; }
            }
        }
        for( int i = hpel_iters; i > 0; i-- )
        {   /* me.c:898-921: candidates in tie-break order up, down, left, right */
            static const int8_t d[4][2] = {{0,-2},{0,2},{-2,0},{2,0}};
            int omx = bmx, omy = bmy, best = -1;
            for( int k = 0; k < 4; k++ )
            {
                int mx = omx + d[k][0], my = omy + d[k][1];
                int cost = cmp_qpel( c, m, mx, my, 0 ) + mvcost( c, m, mx, my );
                if( cost < bcost ) { bcost = cost; best = k; }
            }
            if( best < 0 ) break;
            bmx = omx + d[best][0]; bmy = omy + d[best][1];
        }
    }

    if( c->b_satd )   /* mbcmp != fpelcmp, me.c:925-929 */
        bcost = cmp_qpel( c, m, bmx, bmy, 1 ) + mvcost( c, m, bmx, bmy );

    /* me.c:947-963 (la_subme is 2 or 4, never 1) */
    int bdir = -1;
    for( int i = qpel_iters; i > 0; i-- )
    {
        static const int8_t d[4][2] = {{0,-1},{0,1},{-1,0},{1,0}};
        if( bmy <= m->smin[1] || bmy >= m->smax[1] || bmx <= m->smin[0] || bmx >= m->smax[0] )
            break;
        int odir = bdir, omx = bmx, omy = bmy;
        for( int k = 0; k < 4; k++ )
            if( (k^1) != odir )
            {
                int mx = omx + d[k][0], my = omy + d[k][1];
                int cost = cmp_qpel( c, m, mx, my, 1 ) + mvcost( c, m, mx, my );
                if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; bdir = k; }
            }
        if( bmx == omx && bmy == omy ) break;
    }
    m->cost = bcost; m->mv[0] = bmx; m->mv[1] = bmy;
}

/* x264_me_search_ref, me.c:182-420,774-797, for PIXEL_8x8 in the two lookahead configurations */
static void me_search( xo_ctx *c, me_t *m, int (*mvc)[2], int i_mvc )
{
    int bmx, bmy, bcost = XO_COST_MAX, bpred_cost = XO_COST_MAX;
    int bpred_mx = 0, bpred_my = 0, pmv_x, pmv_y;
    int cand[5][2];
#define COST_MV(mx,my) do { int cost_ = sad_fpel( c, m, mx, my ) + mvcost( c, m, (mx)*4, (my)*4 ); \
                            if( cost_ < bcost ) { bcost = cost_; bmx = mx; bmy = my; } } while(0)

    if( c->la_subme >= 3 )
    {   /* me.c:216-275 */
        bpred_mx = clip3( m->mvp[0], 4*m->fmin[0], 4*m->fmax[0] );
        bpred_my = clip3( m->mvp[1], 4*m->fmin[1], 4*m->fmax[1] );
        pmv_x = bpred_mx; pmv_y = bpred_my;
        bpred_cost = cmp_qpel( c, m, bpred_mx, bpred_my, 0 ) + mvcost( c, m, bpred_mx, bpred_my );
        int pmv_cost = bpred_cost;
        int n = 0;
        for( int i = 0; i < i_mvc; i++ )
        {   /* x264_predictor_clip, common.h:790-805 */
            int mx = mvc[i][0], my = mvc[i][1];
            if( ( !mx && !my ) || ( mx == pmv_x && my == pmv_y ) ) continue;
            cand[n+1][0] = clip3( mx, 4*m->fmin[0], 4*m->fmax[0] );
            cand[n+1][1] = clip3( my, 4*m->fmin[1], 4*m->fmax[1] );
            n++;
        }
        if( n > 0 )
        {
            cand[0][0] = pmv_x; cand[0][1] = pmv_y;
            int best = 0;
            for( int i = 1; i <= n; i++ )
            {
                int cost = cmp_qpel( c, m, cand[i][0], cand[i][1], 0 ) + mvcost( c, m, cand[i][0], cand[i][1] );
                if( cost < bpred_cost ) { bpred_cost = cost; best = i; }
            }
            bpred_mx = cand[best][0]; bpred_my = cand[best][1];
        }
        bmx = FPEL( bpred_mx ); bmy = FPEL( bpred_my );
        if( (bpred_mx&3) | (bpred_my&3) )
            COST_MV( bmx, bmy );
        else
            bcost = bpred_cost;
        if( pmv_x | pmv_y )
        {
            if( bmx | bmy ) COST_MV( 0, 0 );
        }
        else if( pmv_cost < bcost ) { bcost = pmv_cost; bmx = 0; bmy = 0; }
    }
    else
    {   /* me.c:278-317 */
        bmx = clip3( FPEL( m->mvp[0] ), m->fmin[0], m->fmax[0] );
        bmy = clip3( FPEL( m->mvp[1] ), m->fmin[1], m->fmax[1] );
        pmv_x = bmx; pmv_y = bmy;
        bcost = sad_fpel( c, m, bmx, bmy );
        int n = 0;
        for( int i = 0; i < i_mvc; i++ )
        {   /* x264_predictor_roundclip, common.h:774-788 */
            int mx = ( mvc[i][0] + 2 ) >> 2, my = ( mvc[i][1] + 2 ) >> 2;
            if( ( !mx && !my ) || ( mx == pmv_x && my == pmv_y ) ) continue;
            cand[n+1][0] = clip3( mx, m->fmin[0], m->fmax[0] );
            cand[n+1][1] = clip3( my, m->fmin[1], m->fmax[1] );
            n++;
        }
        if( n > 0 )
        {
            cand[0][0] = pmv_x; cand[0][1] = pmv_y;
            int best = 0;
            for( int i = 1; i <= n; i++ )
            {
                int cost = sad_fpel( c, m, cand[i][0], cand[i][1] ) + mvcost( c, m, cand[i][0]*4, cand[i][1]*4 );
                if( cost < bcost ) { bcost = cost; best = i; }
            }
            bmx = cand[best][0]; bmy = cand[best][1];
        }
        if( pmv_x | pmv_y ) COST_MV( 0, 0 );
    }

    if( c->me_method == 0 )
    {   /* DIA, me.c:322-342: tie-break order up, down, left, right */
        static const int8_t d[4][2] = {{0,-1},{0,1},{-1,0},{1,0}};
        int i = c->cfg.me_range;
        do
        {
            int best = -1, ox = bmx, oy = bmy;
            for( int k = 0; k < 4; k++ )
            {
                int cost = sad_fpel( c, m, ox+d[k][0], oy+d[k][1] ) + mvcost( c, m, (ox+d[k][0])*4, (oy+d[k][1])*4 );
                if( cost < bcost ) { bcost = cost; best = k; }
            }
            if( best < 0 ) break;
            bmx = ox + d[best][0]; bmy = oy + d[best][1];
        } while( --i && in_range( m, bmx, bmy ) );
    }
    else
    {   /* HEX, me.c:344-420 */
        static const int8_t first[6] = { 1, 2, 3, 4, 5, 6 };   /* hex2[] rows in code order 2..7: (-2,0),(-1,2),(1,2),(2,0),(1,-2),(-1,-2) */
        int best = -1;
        for( int k = 0; k < 6; k++ )
        {
            int mx = bmx + hex2[first[k]][0], my = bmy + hex2[first[k]][1];
            int cost = sad_fpel( c, m, mx, my ) + mvcost( c, m, mx*4, my*4 );
            if( cost < bcost ) { bcost = cost; best = k; }
        }
        if( best >= 0 )
        {
            int dir = best;                     /* (bcost&7)-2 */
            bmx += hex2[dir+1][0]; bmy += hex2[dir+1][1];
            for( int i = (c->cfg.me_range>>1) - 1; i > 0 && in_range( m, bmx, bmy ); i-- )
            {
                int b2 = -1;
                for( int k = 0; k < 3; k++ )
                {
                    int mx = bmx + hex2[dir+k][0], my = bmy + hex2[dir+k][1];
                    int cost = sad_fpel( c, m, mx, my ) + mvcost( c, m, mx*4, my*4 );
                    if( cost < bcost ) { bcost = cost; b2 = k; }
                }
                if( b2 < 0 ) break;
                dir += b2 - 1;                  /* (bcost&7)-2 with codes 1..3 */
                dir = mod6m1[dir+1];
                bmx += hex2[dir+1][0]; bmy += hex2[dir+1][1];
            }
        }
        /* square refine, me.c:404-418: order (0,-1),(0,1),(-1,0),(1,0),(-1,-1),(-1,1),(1,-1),(1,1) */
        int bs = 0, ox = bmx, oy = bmy;
        for( int k = 1; k <= 8; k++ )
        {
            int mx = ox + square1[k][0], my = oy + square1[k][1];
            int cost = sad_fpel( c, m, mx, my ) + mvcost( c, m, mx*4, my*4 );
            if( cost < bcost ) { bcost = cost; bs = k; }
        }
        bmx = ox + square1[bs][0]; bmy = oy + square1[bs][1];
    }
#undef COST_MV

    /* me.c:774-797 */
    if( c->la_subme < 3 )
    {
        m->cost = bcost;
        if( bmx == pmv_x && bmy == pmv_y ) m->cost += mvcost( c, m, bmx*4, bmy*4 );
        m->mv[0] = bmx*4; m->mv[1] = bmy*4;
        refine_subpel( c, m, 1, 0 );    /* subpel_iterations[2] = {0,1,1,0} */
    }
    else
    {
        if( bpred_cost < bcost ) { m->mv[0] = bpred_mx; m->mv[1] = bpred_my; }
        else                     { m->mv[0] = bmx*4;    m->mv[1] = bmy*4; }
        m->cost = imin( bpred_cost, bcost );
        refine_subpel( c, m, 1, 1 );    /* subpel_iterations[4] = {0,2,1,1} */
    }
}

/* per-evaluation state shared by the block loop */
typedef struct
{
    xo_frame *fenc, *fref[2];
    int p0, p1, b, b_bidir, dist_scale_factor, bipred_weight;
    int do_search[2];
    int wscale, wdenom, woffset;
    int16_t (*mvs[2])[2];
    int *mv_costs[2];
    const int16_t (*dmv_src)[2];     /* fref1->lowres_mvs[0][p1-p0-1] if searched (slicetype.c:629) */
    int final_pass;
    int acc_inter[3], acc_intra[3];  /* COST_EST, COST_EST_AQ, INTRA_MBS */
    int *row_inter, *row_intra;
} eval_t;

static int bidir_cost( xo_ctx *c, eval_t *e, const uint8_t *fenc, const uint8_t *const r0[4], const uint8_t *const r1[4],
                       const int mv0[2], const int mv1[2] )
{   /* TRY_BIDIR, slicetype.c:579-603, without the penalty */
    uint8_t a[64], b[64], avg[64];
    int m0[2] = { mv0[0], mv0[1] }, m1[2] = { mv1[0], mv1[1] };
    if( !c->b_satd )
    {   /* subme <= 1: hpel planes addressed directly, no qpel averaging (slicetype.c:582-589) */
        int i0 = ((m0[0]&2)>>1) + (m0[1]&2), i1 = ((m1[0]&2)>>1) + (m1[1]&2);
        const uint8_t *s0 = r0[i0] + (m0[0]>>2) + (m0[1]>>2)*c->stride;
        const uint8_t *s1 = r1[i1] + (m1[0]>>2) + (m1[1]>>2)*c->stride;
        xo_avg( avg, 8, s0, c->stride, s1, c->stride, 8, 8, e->bipred_weight );
        c->cnt[1]++;
        return xo_sad( XO_8x8, fenc, c->stride, avg, 8 );
    }
    /* both get_ref calls receive the list-0 weight w (slicetype.c:594-597); w is none in B frames */
    xo_mc_luma( a, 8, r0, c->stride, m0[0], m0[1], 8, 8, e->wscale, e->wdenom, e->woffset );
    xo_mc_luma( b, 8, r1, c->stride, m1[0], m1[1], 8, 8, e->wscale, e->wdenom, e->woffset );
    xo_avg( avg, 8, a, 8, b, 8, 8, 8, e->bipred_weight );
    c->cnt[2]++;
    return xo_satd( XO_8x8, fenc, c->stride, avg, 8 );
}

/* slicetype.c:514-791 for block (bx,by). nb: neighbour MV fields of the previous pass. */
static void mb_cost( xo_ctx *c, eval_t *e, int bx, int by, int16_t (*const nb[2])[2] )
{
    const int mbw = c->mbw, mbh = c->mbh, xy = bx + by*mbw, st = c->stride;
    const intptr_t pel = 8 * ( bx + (intptr_t)by*st );
    xo_frame *fenc = e->fenc;
    const uint8_t *src = fenc->plane[0] + pel;
    const int b_score = ( bx > 0 && bx < mbw-1 && by > 0 && by < mbh-1 ) || mbw <= 2 || mbh <= 2;
    int bcost = XO_COST_MAX, list_used = 0;
    const int lowres_penalty = 4, lambda = 1;
    me_t m[2];
    const uint8_t *r[2][4];

    if( e->p0 != e->p1 )
    {
        int mv_range = 2 * c->cfg.mv_range;
        int smin[2], smax[2];
        /* slicetype.c:552-562; the y limits are refreshed at every row start, i.e. they depend on by only */
        smin[0] = imax( 4*(-8*bx - 12), -mv_range );
        smax[0] = imin( 4*(8*(mbw - bx - 1) + 12), mv_range-1 );
        smin[1] = imax( 4*(-8*by - 12), -mv_range );
        smax[1] = imin( 4*(8*(mbh - by - 1) + 12), mv_range-1 );

        for( int l = 0; l < 1 + e->b_bidir; l++ )
        {
            memset( &m[l], 0, sizeof(me_t) );
            m[l].fenc = src; m[l].fs = st; m[l].stride = st;
            for( int k = 0; k < 4; k++ ) m[l].ref[k] = r[l][k] = e->fref[l]->plane[k] + pel;
            m[l].ref_w = m[l].ref[0];
            m[l].wdenom = -1;
            for( int k = 0; k < 2; k++ )
            {
                m[l].smin[k] = smin[k]; m[l].smax[k] = smax[k];
                m[l].fmin[k] = smin[k] >> 2; m[l].fmax[k] = smax[k] >> 2;
            }
        }
        if( e->wdenom >= 0 )
        {   /* slicetype.c:609-614 */
            m[0].wscale = e->wscale; m[0].wdenom = e->wdenom; m[0].woffset = e->woffset;
            m[0].ref_w = e->fref[0]->wplane + pel;
        }

        if( e->b_bidir && e->final_pass )
        {   /* slicetype.c:616-652 */
            int dmv[2][2] = {{0,0},{0,0}};
            if( e->dmv_src )
            {
                const int16_t *mvr = e->dmv_src[xy];
                dmv[0][0] = ( mvr[0] * e->dist_scale_factor + 128 ) >> 8;
                dmv[0][1] = ( mvr[1] * e->dist_scale_factor + 128 ) >> 8;
                dmv[1][0] = dmv[0][0] - mvr[0];
                dmv[1][1] = dmv[0][1] - mvr[1];
                for( int k = 0; k < 2; k++ )
                {
                    dmv[k][0] = clip3( dmv[k][0], smin[0], smax[0] );
                    dmv[k][1] = clip3( dmv[k][1], smin[1], smax[1] );
                    if( !c->b_satd ) { dmv[k][0] &= ~1; dmv[k][1] &= ~1; }
                }
            }
            int cost = bidir_cost( c, e, src, r[0], r[1], dmv[0], dmv[1] );
            if( cost < bcost ) { bcost = cost; list_used = 3; }
            if( dmv[0][0] | dmv[0][1] | dmv[1][0] | dmv[1][1] )
            {
                uint8_t avg[64];
                xo_avg( avg, 8, r[0][0], st, r[1][0], st, 8, 8, e->bipred_weight );
                c->cnt[c->b_satd ? 2 : 1]++;
                cost = c->b_satd ? xo_satd( XO_8x8, src, st, avg, 8 ) : xo_sad( XO_8x8, src, st, avg, 8 );
                if( cost < bcost ) { bcost = cost; list_used = 3; }
            }
        }

        for( int l = 0; l < 1 + e->b_bidir; l++ )
        {
            if( e->do_search[l] )
            {   /* slicetype.c:658-702 */
                int mvc[4][2] = {{0,0},{0,0},{0,0},{0,0}}, i_mvc = 0;
                const int16_t (*nm)[2] = nb[l] + xy;
#define MVC(i) { mvc[i_mvc][0] = nm[i][0]; mvc[i_mvc][1] = nm[i][1]; i_mvc++; }
                if( bx < mbw-1 ) MVC( 1 );
                if( by < mbh-1 )
                {
                    MVC( mbw );
                    if( bx > 0 )     MVC( mbw-1 );
                    if( bx < mbw-1 ) MVC( mbw+1 );
                }
#undef MVC
                if( i_mvc <= 1 ) { m[l].mvp[0] = mvc[0][0]; m[l].mvp[1] = mvc[0][1]; }
                else
                {
                    m[l].mvp[0] = median3( mvc[0][0], mvc[1][0], mvc[2][0] );
                    m[l].mvp[1] = median3( mvc[0][1], mvc[1][1], mvc[2][1] );
                }
                int skip = 0;
                if( !( m[l].mvp[0] | m[l].mvp[1] ) )
                {   /* slicetype.c:684-692: unweighted plane 0 at mv 0 */
                    c->cnt[c->b_satd ? 2 : 1]++;
                    m[l].cost = c->b_satd ? xo_satd( XO_8x8, src, st, m[l].ref[0], st ) : xo_sad( XO_8x8, src, st, m[l].ref[0], st );
                    if( m[l].cost < 64 ) { m[l].mv[0] = m[l].mv[1] = 0; skip = 1; }
                }
                if( !skip )
                {
                    me_search( c, &m[l], mvc, i_mvc );
                    m[l].cost -= c->cost_mv[0];
                    if( m[l].mv[0] | m[l].mv[1] ) m[l].cost += 5*lambda;
                }
                e->mvs[l][xy][0] = m[l].mv[0]; e->mvs[l][xy][1] = m[l].mv[1];
                e->mv_costs[l][xy] = m[l].cost;
            }
            else
            {
                m[l].mv[0] = e->mvs[l][xy][0]; m[l].mv[1] = e->mvs[l][xy][1];
                m[l].cost = e->mv_costs[l][xy];
            }
            if( m[l].cost < bcost ) { bcost = m[l].cost; list_used = l+1; }
        }
        if( !e->final_pass ) return;

        if( e->b_bidir && ( m[0].mv[0] | m[0].mv[1] | m[1].mv[0] | m[1].mv[1] ) )
        {   /* slicetype.c:711-712 */
            int cost = 5*lambda + bidir_cost( c, e, src, r[0], r[1], m[0].mv, m[1].mv );
            if( cost < bcost ) { bcost = cost; list_used = 3; }
        }
    }

    if( !fenc->b_intra_calculated )
    {   /* slicetype.c:715-757 */
        int icost = intra_min_cost( src, st, c->b_satd ) + 5*lambda + lowres_penalty;
        fenc->lowres_costs[0][0][xy] = icost;       /* i_intra_cost aliases lowres_costs[0][0], frame.c:287 */
        int icost_aq = c->cfg.aq ? ( icost * fenc->inv_qscale[xy] + 128 ) >> 8 : icost;
        e->row_intra[by] += icost_aq;
        if( b_score ) { e->acc_intra[0] += icost; e->acc_intra[1] += icost_aq; }
    }
    bcost += lowres_penalty;

    if( !e->b_bidir )
    {   /* slicetype.c:762-773 */
        int icost = fenc->lowres_costs[0][0][xy];
        int b_intra = icost < bcost;
        if( b_intra ) { bcost = icost; list_used = 0; }
        if( b_score ) e->acc_inter[2] += b_intra;
    }
    if( e->p0 != e->p1 )
    {   /* slicetype.c:776-788 */
        int bcost_aq = c->cfg.aq ? ( bcost * fenc->inv_qscale[xy] + 128 ) >> 8 : bcost;
        e->row_inter[by] += bcost_aq;
        if( b_score ) { e->acc_inter[0] += bcost; e->acc_inter[1] += bcost_aq; }
    }
    fenc->lowres_costs[e->b-e->p0][e->p1-e->b][xy] = imin( bcost, 0x3FFF ) + ( list_used << 14 );   /* :790 */
}

static void ensure_arrays( xo_ctx *c, xo_frame *f, int i, int j )
{
    if( !f->lowres_costs[0][0] )
    {
        f->lowres_costs[0][0] = malloc( c->nmb * 2 );
        memset( f->lowres_costs[0][0], 0xFF, c->nmb * 2 );
        f->row_satds[0][0] = calloc( c->mbh, sizeof(int) );
    }
    if( !f->lowres_costs[i][j] ) f->lowres_costs[i][j] = calloc( c->nmb, 2 );
    if( !f->row_satds[i][j] )    f->row_satds[i][j] = calloc( c->mbh, sizeof(int) );
}

/* slicetype_frame_cost + slicetype_slice_cost, slicetype.c:814-995, one slice, dependency-free passes */
int xo_frame_cost( xo_ctx *c, int p0, int p1, int b, int w_scale, int w_denom, int w_offset )
{
    if( p0 < 0 || p1 >= XO_MAX_FRAMES || p0 > b || b > p1 || p1-p0 > c->cfg.bframes+1 ) return -1;
    xo_frame *fenc = &c->fr[b];
    if( !fenc->valid || !c->fr[p0].valid || !c->fr[p1].valid ) return -1;
    memset( c->cnt, 0, sizeof(c->cnt) );
    if( fenc->cost_est[b-p0][p1-b] >= 0 )
        return fenc->cost_est[b-p0][p1-b];

    eval_t e;
    memset( &e, 0, sizeof(e) );
    e.fenc = fenc; e.fref[0] = &c->fr[p0]; e.fref[1] = &c->fr[p1];
    e.p0 = p0; e.p1 = p1; e.b = b; e.b_bidir = b < p1;
    e.wdenom = -1;
    ensure_arrays( c, fenc, b-p0, p1-b );
    for( int l = 0; l < 2; l++ )
    {
        int d = l ? p1-b-1 : b-p0-1;
        if( d < 0 ) continue;
        if( !fenc->lowres_mvs[l][d] )
        {
            fenc->lowres_mvs[l][d] = calloc( c->nmb, 4 );
            fenc->lowres_mv_costs[l][d] = calloc( c->nmb, sizeof(int) );
        }
        e.mvs[l] = fenc->lowres_mvs[l][d];
        e.mv_costs[l] = fenc->lowres_mv_costs[l][d];
        e.do_search[l] = !fenc->searched[l][d];
        fenc->searched[l][d] = 1;
    }
    if( e.do_search[0] && b == p1 && w_denom >= 0 )
    {   /* weights come from the host (x264_weights_analyse stays on the CPU); build weighted[0] */
        xo_frame *ref = &c->fr[p0];
        size_t psz = (size_t)c->stride * c->plane_rows;
        free( ref->wbuf );
        ref->wbuf = malloc( psz );
        const uint8_t *s = ref->plane[0] - XO_PAD*c->stride - XO_PAD;
        for( size_t i = 0; i < psz; i++ ) ref->wbuf[i] = weight_px( s[i], w_scale, w_denom, w_offset );
        ref->wplane = ref->wbuf + XO_PAD*c->stride + XO_PAD;
        e.wscale = w_scale; e.wdenom = w_denom; e.woffset = w_offset;
    }
    e.dist_scale_factor = p1 != p0 ? ( ((b-p0) << 8) + ((p1-p0) >> 1) ) / (p1-p0) : 128;
    e.bipred_weight = c->cfg.weighted_bipred ? 64 - (e.dist_scale_factor>>2) : 32;
    if( e.b_bidir && c->fr[p1].searched[0][p1-p0-1] )
        e.dmv_src = (const int16_t(*)[2])c->fr[p1].lowres_mvs[0][p1-p0-1];
    e.row_inter = fenc->row_satds[b-p0][p1-b];
    e.row_intra = fenc->row_satds[0][0];

    int do_edges = c->cfg.do_edges || c->mbw <= 2 || c->mbh <= 2;
    int start_y = imin( c->mbh-1, c->mbh-2+do_edges ), end_y = imax( 0, 1-do_edges );
    int start_x = c->mbw-2+do_edges, end_x = 1-do_edges;
    const int serial = c->cfg.search_passes == 0;
    int n_pass = ( e.do_search[0] || e.do_search[1] ) && !serial ? c->cfg.search_passes : 1;
    int64_t cnt_total[3] = {0,0,0};

    for( int l = 0; l < 2; l++ )
        if( e.do_search[l] ) memset( e.mvs[l], 0, c->nmb*4 );

    for( int pass = 0; pass < n_pass; pass++ )
    {
        /* serial: neighbours come from the field being written, as in slicetype.c:662-680 (reverse raster order) */
        int16_t (*nb[2])[2] = { serial ? e.mvs[0] : c->snap[0], serial ? e.mvs[1] : c->snap[1] };
        for( int l = 0; l < 2; l++ )
            if( e.do_search[l] && !serial ) memcpy( c->snap[l], e.mvs[l], c->nmb*4 );
        e.final_pass = pass == n_pass-1;
        if( e.final_pass )
        {
            memset( e.acc_inter, 0, sizeof(e.acc_inter) );
            memset( e.acc_intra, 0, sizeof(e.acc_intra) );
            if( p0 != p1 ) memset( e.row_inter, 0, c->mbh*sizeof(int) );
            if( !fenc->b_intra_calculated ) memset( e.row_intra, 0, c->mbh*sizeof(int) );
        }
        for( int by = start_y; by >= end_y; by-- )
            for( int bx = start_x; bx >= end_x; bx-- )
                mb_cost( c, &e, bx, by, nb );
        for( int k = 0; k < 3; k++ ) cnt_total[k] += c->cnt[k];
    }
    memcpy( c->cnt, cnt_total, sizeof(cnt_total) );
    /* note: intermediate passes return before the intra/accumulate section, so only the final pass
     * contributes sums, exactly as if slicetype_mb_cost had been run once with converged neighbours */

    /* slicetype.c:946-989 */
    if( b == p1 ) fenc->intra_mbs[b-p0] = e.acc_inter[2];
    if( !fenc->b_intra_calculated )
    {
        fenc->cost_est[0][0] = e.acc_intra[0];
        fenc->cost_est_aq[0][0] = e.acc_intra[1];
    }
    if( p0 != p1 )
    {
        fenc->cost_est[b-p0][p1-b] = e.acc_inter[0];
        fenc->cost_est_aq[b-p0][p1-b] = e.acc_inter[1];
    }
    int score = fenc->cost_est[b-p0][p1-b];
    if( b != p1 ) score = (int)( (uint64_t)score * 100 / ( 120 + c->cfg.bframe_bias ) );
    else fenc->b_intra_calculated = 1;
    fenc->cost_est[b-p0][p1-b] = score;
    return score;
}

int xo_get( xo_ctx *c, int idx, int what, int i, int j, void *out )
{
    xo_frame *f = &c->fr[idx];
    if( !f->valid ) return -1;
    switch( what )
    {
        case 0: if( !f->lowres_costs[i][j] ) return -1; memcpy( out, f->lowres_costs[i][j], c->nmb*2 ); break;
        case 1: if( !f->lowres_mvs[i][j] ) return -1; memcpy( out, f->lowres_mvs[i][j], c->nmb*4 ); break;
        case 2: if( !f->lowres_mv_costs[i][j] ) return -1; memcpy( out, f->lowres_mv_costs[i][j], c->nmb*4 ); break;
        case 3: if( !f->row_satds[i][j] ) return -1; memcpy( out, f->row_satds[i][j], c->mbh*4 ); break;
        case 4: ((int*)out)[0] = f->cost_est[i][j]; ((int*)out)[1] = f->cost_est_aq[i][j]; ((int*)out)[2] = f->intra_mbs[i]; break;
        case 5: if( !f->lowres_costs[0][0] ) return -1; memcpy( out, f->lowres_costs[0][0], c->nmb*2 ); break;
        case 6: memcpy( out, f->inv_qscale, c->nmb*2 ); break;
        default: return -1;
    }
    return 0;
}

void xo_counters( xo_ctx *c, int64_t *out ) { memcpy( out, c->cnt, sizeof(c->cnt) ); }

/* ================================================================================================ */
/* MB-tree                                                                   */
/* ================================================================================================ */

/* mc.c:511-525 */
void xo_mbtree_propagate_cost( int16_t *dst, const uint16_t *propagate_in, const uint16_t *intra_costs, const uint16_t *inter_costs,
                               const uint16_t *inv_qscales, float fps, int len )
{
    for( int i = 0; i < len; i++ )
    {
        int intra_cost = intra_costs[i];
        int inter_cost = imin( intra_costs[i], inter_costs[i] & 0x3FFF );
        float propagate_intra  = intra_cost * inv_qscales[i];
        float propagate_amount = propagate_in[i] + propagate_intra*fps;
        float propagate_num    = intra_cost - inter_cost;
        float propagate_denom  = intra_cost;
        dst[i] = imin( (int)( propagate_amount * propagate_num / propagate_denom + 0.5f ), 32767 );
    }
}

#define CLIP_ADD(s,x) (s) = imin( (s)+(x), (1<<15)-1 )       /* MC_CLIP_ADD, mc.h:29 */

/* mc.c:527-598 */
void xo_mbtree_propagate_list( int mbw, int mbh, uint16_t *ref_costs, const int16_t (*mvs)[2], const int16_t *propagate_amount,
                               const uint16_t *lowres_costs, int bipred_weight, int mb_y, int len, int list )
{
    unsigned stride = mbw, width = mbw, height = mbh;
    for( int i = 0; i < len; i++ )
    {
        int lists_used = lowres_costs[i] >> 14;
        if( !( lists_used & ( 1 << list ) ) ) continue;
        int listamount = propagate_amount[i];
        if( lists_used == 3 ) listamount = ( listamount * bipred_weight + 32 ) >> 6;
        if( !mvs[i][0] && !mvs[i][1] ) { CLIP_ADD( ref_costs[mb_y*stride + i], listamount ); continue; }
        int x = mvs[i][0], y = mvs[i][1];
        unsigned mbx = (unsigned)( ( x >> 5 ) + i ), mby = (unsigned)( ( y >> 5 ) + mb_y );
        unsigned idx0 = mbx + mby * stride, idx2 = idx0 + stride;
        x &= 31; y &= 31;
        int w0 = ( ( 32-y ) * ( 32-x ) * listamount + 512 ) >> 10, w1 = ( ( 32-y ) * x * listamount + 512 ) >> 10;
        int w2 = ( y * ( 32-x ) * listamount + 512 ) >> 10, w3 = ( y * x * listamount + 512 ) >> 10;
        if( mby < height )
        {
            if( mbx < width ) CLIP_ADD( ref_costs[idx0], w0 );
            if( mbx+1 < width ) CLIP_ADD( ref_costs[idx0+1], w1 );
        }
        if( mby+1 < height )
        {
            if( mbx < width ) CLIP_ADD( ref_costs[idx2], w2 );
            if( mbx+1 < width ) CLIP_ADD( ref_costs[idx2+1], w3 );
        }
    }
}

static uint16_t *prop_of( xo_ctx *c, int idx )
{
    xo_frame *f = &c->fr[idx];
    if( !f->propagate ) f->propagate = calloc( c->nmb, sizeof(uint16_t) );
    return f->propagate;
}

int xo_mbtree_reset( xo_ctx *c, int idx ) { memset( prop_of( c, idx ), 0, c->nmb * sizeof(uint16_t) ); return 0; }
int xo_mbtree_get( xo_ctx *c, int idx, uint16_t *out ) { memcpy( out, prop_of( c, idx ), c->nmb * sizeof(uint16_t) ); return 0; }
int xo_mbtree_swap( xo_ctx *c, int a, int b )
{
    uint16_t *pa = prop_of( c, a ), *pb = prop_of( c, b );
    c->fr[a].propagate = pb; c->fr[b].propagate = pa;
    return 0;
}

/* macroblock_tree_propagate, slicetype.c:1051-1089 (without the VBV-lookahead finish) */
int xo_mbtree_propagate( xo_ctx *c, int p0, int p1, int b, float fps_factor, int referenced )
{
    xo_frame *fb = &c->fr[b];
    if( !fb->valid || !fb->lowres_costs[b-p0][p1-b] ) return -1;
    uint16_t *ref_costs[2] = { prop_of( c, p0 ), prop_of( c, p1 ) };
    int dsf = ( ( ( b-p0 ) << 8 ) + ( ( p1-p0 ) >> 1 ) ) / ( p1-p0 );
    int w = c->cfg.weighted_bipred ? 64 - ( dsf >> 2 ) : 32;
    int bw[2] = { w, 64 - w };
    const int16_t (*mvs[2])[2] = { b != p0 ? (const int16_t(*)[2])fb->lowres_mvs[0][b-p0-1] : NULL,
                                   b != p1 ? (const int16_t(*)[2])fb->lowres_mvs[1][p1-b-1] : NULL };
    uint16_t *propagate_cost = prop_of( c, b );
    const uint16_t *lowres_costs = fb->lowres_costs[b-p0][p1-b];
    int16_t *buf = malloc( c->mbw * sizeof(int16_t) );
    if( !referenced ) memset( propagate_cost, 0, c->mbw * sizeof(uint16_t) );
    for( int y = 0; y < c->mbh; y++ )
    {
        int mb_index = y * c->mbw;
        xo_mbtree_propagate_cost( buf, propagate_cost, fb->lowres_costs[0][0] + mb_index, lowres_costs + mb_index,
                                  fb->inv_qscale + mb_index, fps_factor, c->mbw );
        if( referenced ) propagate_cost += c->mbw;
        xo_mbtree_propagate_list( c->mbw, c->mbh, ref_costs[0], mvs[0] + mb_index, buf, lowres_costs + mb_index, bw[0], y, c->mbw, 0 );
        if( b != p1 )
            xo_mbtree_propagate_list( c->mbw, c->mbh, ref_costs[1], mvs[1] + mb_index, buf, lowres_costs + mb_index, bw[1], y, c->mbw, 1 );
    }
    free( buf );
    return 0;
}

/* x264_log2, base.h:226-230 with the tables of tables.c:66-90 */
static float xo_log2( uint32_t x )
{
    static float lut[128];
    static int init;
    if( !init ) { for( int i = 0; i < 128; i++ ) lut[i] = (float)( round( log2( 1.0 + i / 128.0 ) * 1e5 ) / 1e5 ); init = 1; }
    int lz = __builtin_clz( x );
    return lut[( x << lz >> 24 ) & 0x7f] + (float)( 31 - lz );
}

/* ================================================================================================ */
/* adaptive quantisation and explicit weighted prediction analysis: the producers right before the  */
/* path                                                             */
/* ================================================================================================ */

/* x264_exp2fix8, base.h:216-224; x264_exp2_lut[i] (tables.c:58-64) is round( 256 * ( 2^(i/64) - 1 ) ) */
static int xo_exp2fix8( float x )
{
    static uint8_t lut[64];
    static int init;
    if( !init ) { for( int i = 0; i < 64; i++ ) lut[i] = (uint8_t)lround( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) ); init = 1; }
    int i = x * ( -64.f / 6.f ) + 512.5f;
    if( i < 0 ) return 0;
    if( i > 1023 ) return 0xffff;
    return ( lut[i&63] + 256 ) << ( i >> 6 ) >> 8;
}
void xo_exp2_lut( uint8_t *out64 ) { for( int i = 0; i < 64; i++ ) out64[i] = (uint8_t)lround( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) ); }

/* pixel_var_16x16 / _8x8 (pixel.c:158-175) over a plane with replicated mod-16 padding (frame.c:640-668):
 * step is 1 for luma, 2 for one component of the interleaved NV12 chroma plane */
static uint32_t ac_energy_var( const uint8_t *p, int stride, int x0, int y0, int n, int step, int wmax, int hmax, int shift,
                               uint64_t *acc_sum, uint64_t *acc_ssd )
{
    uint32_t sum = 0, sqr = 0;
    for( int y = 0; y < n; y++ )
        for( int x = 0; x < n; x++ )
        {
            int v = p[(intptr_t)imin( y0+y, hmax ) * stride + imin( x0+x, wmax ) * step];
            sum += v; sqr += v*v;
        }
    *acc_sum += sum; *acc_ssd += sqr;                                  /* ac_energy_var b_store, ratecontrol.c:226-236 */
    return sqr - (uint32_t)( ( (uint64_t)sum * sum ) >> shift );
}

/* x264_adaptive_quant_frame (ratecontrol.c:304-415) for a progressive 4:2:0 frame, aq-mode 0..3 (the auto-variance modes 2
 * and 3 call libm's powf exactly as the reference does, :361), followed by x264_frame_init_lowres through xo_put_frame --
 * the order of encoder.c:3417-3421 */
int xo_put_frame_aq( xo_ctx *c, int idx, const uint8_t *luma, int stride, const uint8_t *chroma, int cstride,
                     int aq_mode, float aq_strength, const float *quant_offsets )
{
    if( aq_mode < 0 || aq_mode > 3 ) return -1;
    const int W = c->cfg.width, H = c->cfg.height;
    const int autovar = aq_mode >= 2 && aq_strength != 0;
    float avg_adj = 0.f, avg_adj_pow2 = 0.f;
    uint64_t sum[3] = {0,0,0}, ssd[3] = {0,0,0};
    float *qp = calloc( c->nmb, sizeof(float) );
    uint16_t *inv = malloc( c->nmb * sizeof(uint16_t) );
    const float strength = aq_strength * 1.0397f;                      /* ratecontrol.c:371 */
    for( int i = 0; i < c->nmb; i++ ) inv[i] = 256;
    for( int my = 0; my < c->mbh; my++ )
        for( int mx = 0; mx < c->mbw; mx++ )
        {   /* ac_energy_mb (ratecontrol.c:262-294): luma 16x16 + both 8x8 chroma blocks */
            uint32_t energy = ac_energy_var( luma, stride, 16*mx, 16*my, 16, 1, W-1, H-1, 8, &sum[0], &ssd[0] );
            energy += ac_energy_var( chroma,   cstride, 8*mx, 8*my, 8, 2, W/2-1, H/2-1, 6, &sum[1], &ssd[1] );
            energy += ac_energy_var( chroma+1, cstride, 8*mx, 8*my, 8, 2, W/2-1, H/2-1, 6, &sum[2], &ssd[2] );
            if( autovar )
            {   /* first pass of the auto-variance modes (ratecontrol.c:355-366) */
                float qp_adj = powf( energy * 1.f + 1, 0.125f );
                qp[mx + my*c->mbw] = qp_adj;
                avg_adj += qp_adj;
                avg_adj_pow2 += qp_adj * qp_adj;
            }
            else if( aq_mode == 1 && aq_strength != 0 )
            {   /* ratecontrol.c:388-399 */
                float qp_adj = strength * ( xo_log2( energy > 1 ? energy : 1 ) - 14.427f );
                if( quant_offsets ) qp_adj += quant_offsets[mx + my*c->mbw];
                qp[mx + my*c->mbw] = qp_adj;
                inv[mx + my*c->mbw] = xo_exp2fix8( qp_adj );
            }
            else if( aq_mode && quant_offsets )
            {   /* ratecontrol.c:318-325 */
                qp[mx + my*c->mbw] = quant_offsets[mx + my*c->mbw];
                inv[mx + my*c->mbw] = xo_exp2fix8( quant_offsets[mx + my*c->mbw] );
            }
        }
    if( autovar )
    {   /* ratecontrol.c:367-396 */
        avg_adj /= c->nmb;
        avg_adj_pow2 /= c->nmb;
        const float strength2 = aq_strength * avg_adj;
        avg_adj = avg_adj - 0.5f * ( avg_adj_pow2 - 14.f ) / avg_adj;
        const float bias_strength = aq_strength;
        for( int i = 0; i < c->nmb; i++ )
        {
            float qp_adj = qp[i];
            if( aq_mode == 3 ) qp_adj = strength2 * ( qp_adj - avg_adj ) + bias_strength * ( 1.f - 14.f / ( qp_adj * qp_adj ) );
            else               qp_adj = strength2 * ( qp_adj - avg_adj );
            if( quant_offsets ) qp_adj += quant_offsets[i];
            qp[i] = qp_adj;
            inv[i] = xo_exp2fix8( qp_adj );
        }
    }
    if( xo_put_frame( c, idx, luma, stride, aq_mode ? inv : NULL ) ) { free( qp ); free( inv ); return -1; }
    xo_frame *f = &c->fr[idx];
    for( int i = 0; i < 3; i++ )
    {   /* remove the mean from the SSD (ratecontrol.c:406-414) */
        uint64_t width = 16*c->mbw >> !!i, height = 16*c->mbh >> !!i;
        f->pixel_sum[i] = sum[i];
        f->pixel_ssd[i] = ssd[i] - ( sum[i] * sum[i] + width * height / 2 ) / ( width * height );
    }
    f->qp_offset_aq = qp;
    f->have_aq = 1;
    free( inv );
    return 0;
}

int xo_get_aq( xo_ctx *c, int idx, float *qp_offset, uint16_t *inv_qscale, uint64_t *sums6 )
{
    xo_frame *f = &c->fr[idx];
    if( !f->valid || !f->have_aq ) return -1;
    if( qp_offset ) memcpy( qp_offset, f->qp_offset_aq, c->nmb * sizeof(float) );
    if( inv_qscale ) memcpy( inv_qscale, f->inv_qscale, c->nmb * sizeof(uint16_t) );
    if( sums6 ) for( int i = 0; i < 3; i++ ) { sums6[2*i] = f->pixel_sum[i]; sums6[2*i+1] = f->pixel_ssd[i]; }
    return 0;
}

/* bs_size_ue / bs_size_se (bitstream.h:278-299); x264_ue_size_tab[i] is 2*floor(log2(i))+1, and 1 for i = 0 */
static int ue_size_tab( int i ) { return i ? 2 * ( 31 - __builtin_clz( i ) ) + 1 : 1; }
static int bs_size_ue( unsigned v ) { return ue_size_tab( v + 1 ); }
static int bs_size_se( int v )
{
    int tmp = 1 - v*2;
    if( tmp < 0 ) tmp = v*2;
    return tmp < 256 ? ue_size_tab( tmp ) : ue_size_tab( tmp >> 8 ) + 16;
}

/* weight_cost_luma (slicetype.c:191-222): src is the reference's plane 0 or its motion-compensated copy (stride 8*mbw) */
static unsigned weight_cost_luma( xo_ctx *c, xo_frame *fenc, const uint8_t *src, int sstride, int weighted, int scale, int denom, int offset, int numslices )
{
    unsigned cost = 0;
    uint8_t buf[64];
    for( int y = 0, i_mb = 0; y < c->lh; y += 8 )
        for( int x = 0; x < c->lw; x += 8, i_mb++ )
        {
            const uint8_t *s = src + (intptr_t)y*sstride + x, *e = fenc->plane[0] + (intptr_t)y*c->stride + x;
            int cmp;
            if( weighted )
            {
                for( int j = 0; j < 8; j++ )
                    for( int i = 0; i < 8; i++ ) buf[8*j+i] = weight_px( s[j*sstride+i], scale, denom, offset );
                cmp = c->b_satd ? xo_satd( XO_8x8, buf, 8, e, c->stride ) : xo_sad( XO_8x8, buf, 8, e, c->stride );
            }
            else
                cmp = c->b_satd ? xo_satd( XO_8x8, s, sstride, e, c->stride ) : xo_sad( XO_8x8, s, sstride, e, c->stride );
            cost += imin( cmp, fenc->lowres_costs[0][0][i_mb] );
        }
    if( weighted )    /* weight_slice_header_cost (slicetype.c:170-189), lambda 1, luma */
        cost += numslices * ( 10 + bs_size_ue( denom ) * 2 + 2 * ( bs_size_se( scale ) + bs_size_se( offset ) ) );
    return cost;
}

/* x264_weights_analyse( h, fenc, ref, b_lookahead = 1 ) (slicetype.c:284-501): luma only, one candidate.
 * out: [0] enabled [1] scale [2] denom [3] offset [4] minscore [5] origscore */
int xo_weights_analyse( xo_ctx *c, int b, int ref, int numslices, int *out )
{
    xo_frame *fenc = &c->fr[b], *fref = &c->fr[ref];
    if( !fenc->valid || !fref->valid || !fenc->have_aq || !fref->have_aq || b <= ref || b-ref > c->cfg.bframes+1 ) return -1;
    const float epsilon = 1.f/128.f;
    out[0] = 0; out[1] = 1; out[2] = 0; out[3] = 0; out[4] = out[5] = 0;
    int zero_bias = !fref->pixel_ssd[0];
    float fenc_var = fenc->pixel_ssd[0] + zero_bias;
    float ref_var  = fref->pixel_ssd[0] + zero_bias;
    float guess_scale = sqrtf( fenc_var / ref_var );
    float fenc_mean = (float)( fenc->pixel_sum[0] + zero_bias ) / ( ( 16*c->mbh ) * ( 16*c->mbw ) ) / ( 1 << 0 );
    float ref_mean  = (float)( fref->pixel_sum[0] + zero_bias ) / ( ( 16*c->mbh ) * ( 16*c->mbw ) ) / ( 1 << 0 );
    if( fabsf( ref_mean - fenc_mean ) < 0.5f && fabsf( 1.f - guess_scale ) < epsilon )
        return 0;                                                      /* early termination, :335-339 */
    /* weight_get_h264( round( guess_scale * 128 ), 0, w ), :62-74 */
    int mindenom = 7, minscale = round( guess_scale * 128 ), minoff = 0, found = 0;
    while( mindenom > 0 && minscale > 127 ) { mindenom--; minscale >>= 1; }
    minscale = imin( minscale, 127 );
    if( !fenc->b_intra_calculated ) xo_frame_cost( c, b, b, b, 0, -1, 0 );     /* :361-366 */
    /* weight_cost_init_luma (:76-103): motion-compensate the reference when the field has been searched */
    const uint8_t *mcbuf = fref->plane[0]; int mcstride = c->stride;
    uint8_t *tmp = NULL;
    if( fenc->searched[0][b-ref-1] )
    {
        tmp = malloc( (size_t)c->lw * c->lh );
        for( int y = 0, i = 0; y < c->lh; y += 8 )
            for( int x = 0; x < c->lw; x += 8, i++ )
            {
                const uint8_t *pl[4];
                for( int k = 0; k < 4; k++ ) pl[k] = fref->plane[k] + (intptr_t)y*c->stride + x;
                xo_mc_luma( tmp + (size_t)y*c->lw + x, c->lw, pl, c->stride, fenc->lowres_mvs[0][b-ref-1][i][0], fenc->lowres_mvs[0][b-ref-1][i][1], 8, 8, 1, -1, 0 );
            }
        mcbuf = tmp; mcstride = c->lw;
    }
    unsigned origscore, minscore;
    origscore = minscore = weight_cost_luma( c, fenc, mcbuf, mcstride, 0, 0, 0, 0, numslices );
    out[4] = out[5] = origscore;
    if( !minscore ) { free( tmp ); return 0; }
    {   /* the lookahead checks one scale and one offset (scale_dist = offset_dist = 0), :399-441 */
        int cur_scale = minscale;
        int cur_offset = fenc_mean - ref_mean * cur_scale / ( 1 << mindenom ) + 0.5f * 1;
        if( cur_offset < -128 || cur_offset > 127 )
        {
            cur_offset = cur_offset < -128 ? -128 : 127;
            float v = ( 1 << mindenom ) * ( fenc_mean - cur_offset ) / ref_mean + 0.5f;
            cur_scale = v < 0 ? 0 : v > 127 ? 127 : v;
        }
        int i_off = cur_offset < -128 ? -128 : cur_offset > 127 ? 127 : cur_offset;
        unsigned s = weight_cost_luma( c, fenc, mcbuf, mcstride, 1, cur_scale, mindenom, i_off, numslices );
        if( s < minscore ) { minscore = s; minscale = cur_scale; minoff = i_off; found = 1; }
    }
    free( tmp );
    while( mindenom > 0 && !( minscale & 1 ) ) { mindenom--; minscale >>= 1; }       /* :444-451 */
    out[4] = minscore;
    if( !found || ( minscale == 1 << mindenom && minoff == 0 ) || (float)minscore / origscore > 0.998f )
        return 0;
    out[0] = 1; out[1] = minscale; out[2] = mindenom; out[3] = minoff;
    return 0;
}

/* macroblock_tree_finish, slicetype.c:1029-1049 */
int xo_mbtree_finish( xo_ctx *c, int idx, int fps_factor, float weightdelta, float strength, const float *qp_offset_aq, float *qp_offset )
{
    xo_frame *f = &c->fr[idx];
    const uint16_t *prop = prop_of( c, idx );
    for( int i = 0; i < c->nmb; i++ )
    {
        int intra_cost = ( f->lowres_costs[0][0][i] * f->inv_qscale[i] + 128 ) >> 8;
        if( intra_cost )
        {
            int propagate_cost = ( prop[i] * fps_factor + 128 ) >> 8;
            float log2_ratio = xo_log2( intra_cost + propagate_cost ) - xo_log2( intra_cost ) + weightdelta;
            qp_offset[i] = qp_offset_aq[i] - strength * log2_ratio;
        }
    }
    return 0;
}
