// primitives.cu -- batched pixel / dct / quant primitives (the x264_pixel_function_t,
// x264_dct_function_t and x264_quant_function_t entries named by the north star), sm_100a.
//
// Two shapes per primitive:
//   *_offsets : n independent block pairs at arbitrary byte offsets (checkasm's pbuf2+j pattern,
//               tools/checkasm.c:393-397); one thread per block pair, alignment-agnostic loads.
//   *_grid    : every non-overlapping block of a (tall) image pair; one thread per 8-pixel-wide
//               column of a block row, so a warp reads 256 contiguous bytes per image row.
// The arithmetic is the reference's packed 16:16 form (pixel.c:242-259), word for word.
#include "primitives.cuh"
#include <type_traits>

__constant__ int c_bw[8] = { 16, 16, 8, 8, 8, 4, 4, 4 };
__constant__ int c_bh[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };
static const int h_bw[8] = { 16, 16, 8, 8, 8, 4, 4, 4 };
static const int h_bh[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

__global__ void fill_u16_kernel( uint16_t *dst, int value, int n )
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n ) dst[i] = (uint16_t)value;
}
void launch_fill_u16( uint16_t *dst, int value, int n, cudaStream_t st )
{
    fill_u16_kernel<<<( n + 255 ) / 256, 256, 0, st>>>( dst, value, n );
}

// ---- per-thread kernels on rows held as words -------------------------------------------------
struct Ld8  { __device__ __forceinline__ uint2 operator()( const uint8_t *p ) const { return load8u( p ); } };
struct Ld8A { __device__ __forceinline__ uint2 operator()( const uint8_t *p ) const { return __ldg( (const uint2 *)p ); } };

// packed sum of |coef| of one 8x4 (pixel.c:290-307); satd_8x4 = (lo+hi)>>1
template<class LD1, class LD2>
__device__ __forceinline__ u32 satd_8x4_sum( const uint8_t *p1, int s1, const uint8_t *p2, int s2, LD1 ld1, LD2 ld2 )
{
    u32 t[4][4];
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        u32 d[4];
        diff_row8( ld1( p1 + (size_t)i*s1 ), ld2( p2 + (size_t)i*s2 ), d );
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
    return sum;
}
__device__ __forceinline__ int fold16( u32 s ) { return (int)( ( s & 0xffff ) + ( s >> 16 ) ); }

// pixel.c:265-288
__device__ __forceinline__ int satd_4x4( u32 a[4], u32 b[4] )   // one word (4 px) per row
{
    u32 t0[4], t1[4];
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        u32 a0 = byte_of( a[i], 0 ) - byte_of( b[i], 0 ), a1 = byte_of( a[i], 1 ) - byte_of( b[i], 1 );
        u32 a2 = byte_of( a[i], 2 ) - byte_of( b[i], 2 ), a3 = byte_of( a[i], 3 ) - byte_of( b[i], 3 );
        u32 b0 = ( a0 + a1 ) + ( ( a0 - a1 ) << 16 ), b1 = ( a2 + a3 ) + ( ( a2 - a3 ) << 16 );
        t0[i] = b0 + b1; t1[i] = b0 - b1;
    }
    u32 a0, a1, a2, a3, sum = 0;
    HADAMARD4( a0, a1, a2, a3, t0[0], t0[1], t0[2], t0[3] );
    a0 = abs2( a0 ) + abs2( a1 ) + abs2( a2 ) + abs2( a3 ); sum += ( a0 & 0xffff ) + ( a0 >> 16 );
    HADAMARD4( a0, a1, a2, a3, t1[0], t1[1], t1[2], t1[3] );
    a0 = abs2( a0 ) + abs2( a1 ) + abs2( a2 ) + abs2( a3 ); sum += ( a0 & 0xffff ) + ( a0 >> 16 );
    return (int)( sum >> 1 );
}

// pixel.c:334-366.  The sum of |coef| over all 64 coefficients does not depend on the order of the butterfly stages,
// so neighbouring columns are packed into one word (p[2k] | p[2k+1] << 16: two PRMTs per input word), the packed
// words go through the other two horizontal stages and the 8-point vertical transform, and the remaining horizontal
// stage -- the butterfly BETWEEN the two halves of each word -- is folded into the absolute sum:
// |lo+hi| + |lo-hi| == 2*max(|lo|,|hi|).  Fields stay below 2^15 (8 * 4 * 255 = 8160).
__device__ __forceinline__ void diff_row8_pairs( uint2 a, uint2 b, u32 d[4] )
{
    d[0] = __byte_perm( a.x, 0, 0x4140 ) - __byte_perm( b.x, 0, 0x4140 );
    d[1] = __byte_perm( a.x, 0, 0x4342 ) - __byte_perm( b.x, 0, 0x4342 );
    d[2] = __byte_perm( a.y, 0, 0x4140 ) - __byte_perm( b.y, 0, 0x4140 );
    d[3] = __byte_perm( a.y, 0, 0x4342 ) - __byte_perm( b.y, 0, 0x4342 );
}
template<class LD1, class LD2>
__device__ __forceinline__ int sa8d_8x8_raw( const uint8_t *p1, int s1, const uint8_t *p2, int s2, LD1 ld1, LD2 ld2 )
{
    u32 t[8][4];
#pragma unroll
    for( int i = 0; i < 8; i++ )
    {
        u32 d[4];
        diff_row8_pairs( ld1( p1 + (size_t)i*s1 ), ld2( p2 + (size_t)i*s2 ), d );
        HADAMARD4( t[i][0], t[i][1], t[i][2], t[i][3], d[0], d[1], d[2], d[3] );
    }
    u32 sum = 0;
#pragma unroll
    for( int c = 0; c < 4; c++ )
    {
        u32 a[8];
        HADAMARD4( a[0], a[1], a[2], a[3], t[0][c], t[1][c], t[2][c], t[3][c] );
        HADAMARD4( a[4], a[5], a[6], a[7], t[4][c], t[5][c], t[6][c], t[7][c] );
        u32 col = 0;      // per half: 4 maxima <= 4 * 8160 < 2^15
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            const u32 u = abs2( a[k] + a[k+4] ), v = abs2( a[k] - a[k+4] );
            // (lo_u, lo_v) against (hi_u, hi_v): both maxima from one packed max
            col += __vmaxu2( __byte_perm( u, v, 0x5410 ), __byte_perm( u, v, 0x7632 ) );
        }
        sum += ( col & 0xffff ) + ( col >> 16 );
    }
    return (int)( 2 * sum );
}

// pixel.c:383-418: returns (sum8<<32)+sum4 of one 8x8.  sum4 / sum8 are the sums of |coefficient| of the four 4x4 and of the
// 8x8 Hadamard transform of the SOURCE block (minus DC), and a sum of absolute values does not depend on the order of the
// butterfly stages.  Columns k and k+4 share a word (expand_row8, the layout of x264_pixel_satd_8x4), so one HADAMARD4 per row
// is the 4-point horizontal transform of the left (low halves) and right (high halves) 4x4 blocks at once and one per column
// word the vertical one of the top / bottom blocks: sum4 is the packed abs-sum of those 32 words.  For sum8 the butterfly
// between top and bottom is explicit and the one between the halves of a word is folded into the absolute sum,
// |lo+hi| + |lo-hi| == 2*max(|lo|,|hi|), two words per packed max.  Every 16-bit field stays in range for 8-bit pixels:
// |coefficient| <= 64*255, a packed accumulator half <= 65280.  ~500 integer instructions per 8x8 (was ~600 byte by byte).
template<class LD>
__device__ __forceinline__ uint64_t hadamard_ac_8x8( const uint8_t *p, int st, LD ld )
{
    u32 v[8][4];
#pragma unroll
    for( int i = 0; i < 8; i++ )
    {
        u32 e[4];
        expand_row8( ld( p + (size_t)i*st ), e );
        HADAMARD4( v[i][0], v[i][1], v[i][2], v[i][3], e[0], e[1], e[2], e[3] );
    }
    u32 sum4 = 0, acc8 = 0, dc = 0;
#pragma unroll
    for( int c = 0; c < 4; c++ )
    {
        u32 t[4], b[4];
        HADAMARD4( t[0], t[1], t[2], t[3], v[0][c], v[1][c], v[2][c], v[3][c] );
        HADAMARD4( b[0], b[1], b[2], b[3], v[4][c], v[5][c], v[6][c], v[7][c] );
        if( c == 0 ) dc = t[0] + b[0];                      // the four 4x4 DC terms, two per word
#pragma unroll
        for( int j = 0; j < 4; j++ )
        {
            sum4 += abs2( t[j] ) + abs2( b[j] );
            const u32 u = abs2( t[j] + b[j] ), d = abs2( t[j] - b[j] );
            acc8 += __vmaxu2( __byte_perm( u, d, 0x5410 ), __byte_perm( u, d, 0x7632 ) );
        }
    }
    dc = ( dc & 0xffff ) + ( dc >> 16 );
    const u32 s4 = ( sum4 & 0xffff ) + ( sum4 >> 16 ) - dc;
    const u32 s8 = 2 * ( ( acc8 & 0xffff ) + ( acc8 >> 16 ) ) - dc;
    return ( (uint64_t)s8 << 32 ) + s4;
}

// dct.c:157-189: one 4x4, rows as words; writes 16 coefficients
__device__ __forceinline__ void sub4x4_dct( int16_t *dct, const u32 a[4], const u32 b[4] )
{
    int tmp[16];
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int d0 = (int)byte_of( a[i], 0 ) - (int)byte_of( b[i], 0 ), d1 = (int)byte_of( a[i], 1 ) - (int)byte_of( b[i], 1 );
        int d2 = (int)byte_of( a[i], 2 ) - (int)byte_of( b[i], 2 ), d3 = (int)byte_of( a[i], 3 ) - (int)byte_of( b[i], 3 );
        int s03 = d0 + d3, s12 = d1 + d2, d03 = d0 - d3, d12 = d1 - d2;
        tmp[0*4+i] = s03 + s12; tmp[1*4+i] = 2*d03 + d12; tmp[2*4+i] = s03 - s12; tmp[3*4+i] = d03 - 2*d12;
    }
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int s03 = tmp[i*4+0] + tmp[i*4+3], s12 = tmp[i*4+1] + tmp[i*4+2];
        int d03 = tmp[i*4+0] - tmp[i*4+3], d12 = tmp[i*4+1] - tmp[i*4+2];
        dct[i*4+0] = (int16_t)( s03 + s12 ); dct[i*4+1] = (int16_t)( 2*d03 + d12 );
        dct[i*4+2] = (int16_t)( s03 - s12 ); dct[i*4+3] = (int16_t)( d03 - 2*d12 );
    }
}

// ---- metric of a whole block, one thread -------------------------------------------------------
template<class LD1, class LD2>
__device__ __forceinline__ int block_metric( int metric, int bw, int bh, const uint8_t *p1, int s1, const uint8_t *p2, int s2, LD1 ld1, LD2 ld2 )
{
    int sum = 0;
    if( metric == 0 )
    {   // pixel.c:55-80
        for( int y = 0; y < bh; y++ )
        {
            if( bw == 4 ) sum += __vsadu4( load4u( p1 + (size_t)y*s1 ), load4u( p2 + (size_t)y*s2 ) );
            else
                for( int x = 0; x < bw; x += 8 )
                    sum += sad8( ld1( p1 + (size_t)y*s1 + x ), ld2( p2 + (size_t)y*s2 + x ) );
        }
    }
    else if( metric == 1 )
    {   // pixel.c:311-332
        for( int y = 0; y < bh; y += 4 )
        {
            if( bw == 4 )
            {
                u32 a[4], b[4];
#pragma unroll
                for( int i = 0; i < 4; i++ ) { a[i] = load4u( p1 + (size_t)( y+i )*s1 ); b[i] = load4u( p2 + (size_t)( y+i )*s2 ); }
                sum += satd_4x4( a, b );
            }
            else
                for( int x = 0; x < bw; x += 8 )
                    sum += fold16( satd_8x4_sum( p1 + (size_t)y*s1 + x, s1, p2 + (size_t)y*s2 + x, s2, ld1, ld2 ) ) >> 1;
        }
    }
    else
    {   // pixel.c:368-381
        for( int y = 0; y < bh; y += 8 )
            for( int x = 0; x < bw; x += 8 )
                sum += sa8d_8x8_raw( p1 + (size_t)y*s1 + x, s1, p2 + (size_t)y*s2 + x, s2, ld1, ld2 );
        sum = ( sum + 2 ) >> 2;
    }
    return sum;
}

__global__ void __launch_bounds__(128) pixel_cmp_offsets_kernel( int metric, int size, const uint8_t *__restrict__ p1, int s1,
        const uint8_t *__restrict__ p2, int s2, int64_t n, const int64_t *__restrict__ off1, const int64_t *__restrict__ off2, int32_t *__restrict__ out )
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n ) return;
    out[i] = block_metric( metric, c_bw[size], c_bh[size], p1 + off1[i], s1, p2 + off2[i], s2, Ld8(), Ld8() );
}

int launch_pixel_cmp( int metric, int size, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n,
                      const int64_t *off1, const int64_t *off2, int32_t *out, cudaStream_t st )
{
    if( metric < 0 || metric > 2 || size < 0 || size > 7 ) return -1;
    if( metric == 2 && size != 0 && size != 3 ) return -1;            // sa8d: 16x16 and 8x8 only (pixel.c:855-856)
    if( n <= 0 ) return 0;
    pixel_cmp_offsets_kernel<<<(unsigned)( ( n + 127 ) / 128 ), 128, 0, st>>>( metric, size, p1, s1, p2, s2, n, off1, off2, out );
    return 0;
}

// ---- grid shape ---------------------------------------------------------------------------------
// One thread per 8-pixel column of one block row.  BW in {4,8,16}: a thread covers two 4-wide
// blocks, one 8-wide block, or half of a 16-wide block (halves are combined with one shuffle).
template<int METRIC, int BW, int BH, bool ALIGNED2>
__global__ void __launch_bounds__(128) pixel_cmp_grid_kernel( const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2,
                                                              int cols8, int nbx, int32_t *__restrict__ out )
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // 8-px column index
    const int by = blockIdx.y;
    const bool live = c < cols8;
    const uint8_t *a = p1 + (size_t)by * BH * s1 + (size_t)min( c, cols8-1 ) * 8;
    const uint8_t *b = p2 + (size_t)by * BH * s2 + (size_t)min( c, cols8-1 ) * 8;
    typedef typename std::conditional<ALIGNED2, Ld8A, Ld8>::type LD2;
    int v0 = 0, v1 = 0;
    if( METRIC == 0 )
    {
#pragma unroll 4
        for( int y = 0; y < BH; y++ )
        {
            uint2 ra = Ld8A()( a + (size_t)y*s1 ), rb = LD2()( b + (size_t)y*s2 );
            v0 += __vsadu4( ra.x, rb.x ); v1 += __vsadu4( ra.y, rb.y );
        }
        if( BW != 4 ) v0 += v1;
    }
    else if( METRIC == 1 )
    {
        if( BW == 4 )
        {
            for( int y = 0; y < BH; y += 4 )
            {
                u32 al[4], ah[4], bl[4], bh_[4];
#pragma unroll
                for( int i = 0; i < 4; i++ )
                {
                    uint2 ra = Ld8A()( a + (size_t)( y+i )*s1 ), rb = LD2()( b + (size_t)( y+i )*s2 );
                    al[i] = ra.x; ah[i] = ra.y; bl[i] = rb.x; bh_[i] = rb.y;
                }
                v0 += satd_4x4( al, bl ); v1 += satd_4x4( ah, bh_ );
            }
        }
        else
            for( int y = 0; y < BH; y += 4 )
                v0 += fold16( satd_8x4_sum( a + (size_t)y*s1, s1, b + (size_t)y*s2, s2, Ld8A(), LD2() ) ) >> 1;
    }
    else
    {
        for( int y = 0; y < BH; y += 8 )
            v0 += sa8d_8x8_raw( a + (size_t)y*s1, s1, b + (size_t)y*s2, s2, Ld8A(), LD2() );
    }
    if( BW == 16 )
    {
        v0 += __shfl_xor_sync( 0xffffffffu, v0, 1 );
        if( METRIC == 2 ) v0 = ( v0 + 2 ) >> 2;
        if( live && !( c & 1 ) ) out[(size_t)by * nbx + ( c >> 1 )] = v0;
    }
    else if( BW == 8 )
    {
        if( METRIC == 2 ) v0 = ( v0 + 2 ) >> 2;
        if( live ) out[(size_t)by * nbx + c] = v0;
    }
    else if( live )
    {
        int2 *o = (int2 *)( out + (size_t)by * nbx + 2*c );      // nbx even for BW 4 when width % 8 == 0
        *o = make_int2( v0, v1 );
    }
}

template<int METRIC, int BW, int BH>
static int cmp_grid_launch( const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height, int32_t *out, cudaStream_t st )
{
    int cols8 = width / 8, nbx = width / BW, nby = height / BH;
    if( cols8 <= 0 || nby <= 0 ) return 0;
    dim3 grid( ( cols8 + 127 ) / 128, nby );
    bool al2 = ( ( (uintptr_t)p2 | (uintptr_t)s2 ) & 7 ) == 0;
    if( al2 ) pixel_cmp_grid_kernel<METRIC, BW, BH, true><<<grid, 128, 0, st>>>( p1, s1, p2, s2, cols8, nbx, out );
    else      pixel_cmp_grid_kernel<METRIC, BW, BH, false><<<grid, 128, 0, st>>>( p1, s1, p2, s2, cols8, nbx, out );
    return 0;
}

int launch_pixel_cmp_grid( int metric, int size, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height,
                           int32_t *out, cudaStream_t st )
{
    if( ( ( (uintptr_t)p1 | (uintptr_t)s1 ) & 7 ) || ( width & 7 ) ) return -2;      // pix1 rows must be 8-byte aligned
#define CASE( M, S, W, H ) if( metric == M && size == S ) return cmp_grid_launch<M, W, H>( p1, s1, p2, s2, width, height, out, st );
    CASE( 0, 0, 16, 16 ) CASE( 0, 1, 16, 8 ) CASE( 0, 2, 8, 16 ) CASE( 0, 3, 8, 8 ) CASE( 0, 4, 8, 4 ) CASE( 0, 5, 4, 8 ) CASE( 0, 6, 4, 4 ) CASE( 0, 7, 4, 16 )
    CASE( 1, 0, 16, 16 ) CASE( 1, 1, 16, 8 ) CASE( 1, 2, 8, 16 ) CASE( 1, 3, 8, 8 ) CASE( 1, 4, 8, 4 ) CASE( 1, 5, 4, 8 ) CASE( 1, 6, 4, 4 ) CASE( 1, 7, 4, 16 )
    CASE( 2, 0, 16, 16 ) CASE( 2, 3, 8, 8 )
#undef CASE
    return -1;
}

// ---- hadamard_ac ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t hadamard_ac_finish( uint64_t sum )
{   // pixel.c:430
    return ( ( sum >> 34 ) << 32 ) + ( (uint32_t)sum >> 1 );
}

__global__ void __launch_bounds__(128) hadamard_ac_offsets_kernel( int size, const uint8_t *__restrict__ p, int st, int64_t n,
                                                                   const int64_t *__restrict__ off, uint64_t *__restrict__ out )
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n ) return;
    const uint8_t *q = p + off[i];
    uint64_t sum = 0;
    for( int y = 0; y < c_bh[size]; y += 8 )
        for( int x = 0; x < c_bw[size]; x += 8 )
            sum += hadamard_ac_8x8( q + (size_t)y*st + x, st, Ld8() );
    out[i] = hadamard_ac_finish( sum );
}

int launch_hadamard_ac( int size, const uint8_t *p, int stride, int64_t n, const int64_t *off, uint64_t *out, cudaStream_t st )
{
    if( size < 0 || size > 3 ) return -1;
    if( n <= 0 ) return 0;
    hadamard_ac_offsets_kernel<<<(unsigned)( ( n + 127 ) / 128 ), 128, 0, st>>>( size, p, stride, n, off, out );
    return 0;
}

template<int BW, int BH>
__global__ void __launch_bounds__(128) hadamard_ac_grid_kernel( const uint8_t *__restrict__ p, int s, int cols8, int nbx, uint64_t *__restrict__ out )
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int by = blockIdx.y;
    const bool live = c < cols8;
    const uint8_t *a = p + (size_t)by * BH * s + (size_t)min( c, cols8-1 ) * 8;
    uint64_t sum = 0;
    for( int y = 0; y < BH; y += 8 ) sum += hadamard_ac_8x8( a + (size_t)y*s, s, Ld8A() );
    if( BW == 16 )
    {
        sum += __shfl_xor_sync( 0xffffffffu, sum, 1 );
        if( live && !( c & 1 ) ) out[(size_t)by * nbx + ( c >> 1 )] = hadamard_ac_finish( sum );
    }
    else if( live ) out[(size_t)by * nbx + c] = hadamard_ac_finish( sum );
}

int launch_hadamard_ac_grid( int size, const uint8_t *p, int stride, int width, int height, uint64_t *out, cudaStream_t st )
{
    if( size < 0 || size > 3 ) return -1;
    if( ( ( (uintptr_t)p | (uintptr_t)stride ) & 7 ) || ( width & 7 ) ) return -2;
    int cols8 = width / 8, nbx = width / h_bw[size], nby = height / h_bh[size];
    if( cols8 <= 0 || nby <= 0 ) return 0;
    dim3 grid( ( cols8 + 127 ) / 128, nby );
    switch( size )
    {
        case 0: hadamard_ac_grid_kernel<16,16><<<grid, 128, 0, st>>>( p, stride, cols8, nbx, out ); break;
        case 1: hadamard_ac_grid_kernel<16,8><<<grid, 128, 0, st>>>( p, stride, cols8, nbx, out ); break;
        case 2: hadamard_ac_grid_kernel<8,16><<<grid, 128, 0, st>>>( p, stride, cols8, nbx, out ); break;
        case 3: hadamard_ac_grid_kernel<8,8><<<grid, 128, 0, st>>>( p, stride, cols8, nbx, out ); break;
    }
    return 0;
}

// ---- sub_dct --------------------------------------------------------------------------------------
// one thread per 4x4; 4x4 blocks of a bsize block are emitted in the reference's nested Z order (dct.c:191-205)
__device__ __forceinline__ void z_order( int bsize, int k, int &ox, int &oy )
{   // k-th 4x4 of a block: bits (b0: x of 4x4 in 8x8, b1: y), (b2: x of 8x8 in 16x16, b3: y)
    ox = ( ( k & 1 ) + ( ( k >> 2 ) & 1 ) * 2 ) * 4;
    oy = ( ( ( k >> 1 ) & 1 ) + ( ( k >> 3 ) & 1 ) * 2 ) * 4;
    (void)bsize;
}

__global__ void __launch_bounds__(128) sub_dct_offsets_kernel( int bsize, const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2,
        int64_t n, const int64_t *__restrict__ off1, const int64_t *__restrict__ off2, int16_t *__restrict__ out )
{
    const int per = ( bsize / 4 ) * ( bsize / 4 );
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( t >= n * per ) return;
    int64_t i = t / per; int k = (int)( t % per ), ox, oy;
    z_order( bsize, k, ox, oy );
    const uint8_t *a = p1 + off1[i] + (size_t)oy*s1 + ox, *b = p2 + off2[i] + (size_t)oy*s2 + ox;
    u32 ra[4], rb[4];
#pragma unroll
    for( int y = 0; y < 4; y++ ) { ra[y] = load4u( a + (size_t)y*s1 ); rb[y] = load4u( b + (size_t)y*s2 ); }
    __align__(16) int16_t d[16];
    sub4x4_dct( d, ra, rb );
    uint4 *o = (uint4 *)( out + t*16 );
    o[0] = ( (uint4 *)d )[0]; o[1] = ( (uint4 *)d )[1];
}

int launch_sub_dct( int bsize, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n, const int64_t *off1,
                    const int64_t *off2, int16_t *out, cudaStream_t st )
{
    if( bsize != 4 && bsize != 8 && bsize != 16 ) return -1;
    if( n <= 0 ) return 0;
    int64_t nt = n * ( bsize/4 ) * ( bsize/4 );
    sub_dct_offsets_kernel<<<(unsigned)( ( nt + 127 ) / 128 ), 128, 0, st>>>( bsize, p1, s1, p2, s2, n, off1, off2, out );
    return 0;
}

// grid: thread per 4x4 ordered so that consecutive threads write consecutive 32-byte outputs
__global__ void __launch_bounds__(128) sub_dct_grid_kernel( int bsize, const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2,
                                                            int nbx, int64_t nt, int16_t *__restrict__ out )
{
    const int per = ( bsize / 4 ) * ( bsize / 4 );
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( t >= nt ) return;
    int64_t blk = t / per; int k = (int)( t % per ), ox, oy;
    z_order( bsize, k, ox, oy );
    int64_t by = blk / nbx; int bx = (int)( blk % nbx );
    const uint8_t *a = p1 + ( by*bsize + oy ) * s1 + bx*bsize + ox, *b = p2 + ( by*bsize + oy ) * s2 + bx*bsize + ox;
    u32 ra[4], rb[4];
#pragma unroll
    for( int y = 0; y < 4; y++ ) { ra[y] = __ldg( (const u32 *)( a + (size_t)y*s1 ) ); rb[y] = __ldg( (const u32 *)( b + (size_t)y*s2 ) ); }
    __align__(16) int16_t d[16];
    sub4x4_dct( d, ra, rb );
    uint4 *o = (uint4 *)( out + t*16 );
    o[0] = ( (uint4 *)d )[0]; o[1] = ( (uint4 *)d )[1];
}

int launch_sub_dct_grid( int bsize, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height, int16_t *out, cudaStream_t st )
{
    if( bsize != 4 && bsize != 8 && bsize != 16 ) return -1;
    if( ( ( (uintptr_t)p1 | (uintptr_t)s1 | (uintptr_t)p2 | (uintptr_t)s2 ) & 3 ) ) return -2;
    int nbx = width / bsize, nby = height / bsize;
    int64_t nt = (int64_t)nbx * nby * ( bsize/4 ) * ( bsize/4 );
    if( nt <= 0 ) return 0;
    sub_dct_grid_kernel<<<(unsigned)( ( nt + 127 ) / 128 ), 128, 0, st>>>( bsize, p1, s1, p2, s2, nbx, nt, out );
    return 0;
}

// ---- quant ---------------------------------------------------------------------------------------
// quant.c:50-73; 8 coefficients (one 128-bit word) per thread, nz reduced over the block's threads
template<int N>
__global__ void __launch_bounds__(256) quant_kernel( int16_t *__restrict__ coefs, int64_t nblocks, const uint16_t *__restrict__ mf,
                                                     const uint16_t *__restrict__ bias, int32_t *__restrict__ nz )
{
    __shared__ uint16_t s_mf[N], s_bias[N];
    if( threadIdx.x < N ) { s_mf[threadIdx.x] = mf[threadIdx.x]; s_bias[threadIdx.x] = bias[threadIdx.x]; }
    __syncthreads();
    constexpr int TPB = N / 8;                      // threads per block of coefficients
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < nblocks * TPB;
    int k0 = (int)( t % TPB ) * 8, any = 0;
    if( live )
    {
        uint4 v = ( (uint4 *)coefs )[t];
        u32 w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for( int j = 0; j < 4; j++ )
        {
            u32 r = 0;
#pragma unroll
            for( int h = 0; h < 2; h++ )
            {
                int c = (int16_t)( w[j] >> ( 16*h ) );
                int k = k0 + 2*j + h;
                int q = c > 0 ? (int)( ( (u32)s_bias[k] + (u32)c ) * s_mf[k] >> 16 )
                              : -(int)( ( (u32)s_bias[k] + (u32)( -c ) ) * s_mf[k] >> 16 );
                any |= q;
                r |= ( (u32)q & 0xffffu ) << ( 16*h );
            }
            w[j] = r;
        }
        ( (uint4 *)coefs )[t] = make_uint4( w[0], w[1], w[2], w[3] );
    }
    any = ( any & 0xffff ) != 0;                    // nz |= (coef) on int16 values
#pragma unroll
    for( int m = 1; m < TPB; m <<= 1 ) any |= __shfl_xor_sync( 0xffffffffu, any, m );
    if( live && k0 == 0 ) nz[t / TPB] = any;
}

int launch_quant( int b8x8, int16_t *coefs, int64_t nblocks, const uint16_t *mf, const uint16_t *bias, int32_t *nz, cudaStream_t st )
{
    if( nblocks <= 0 ) return 0;
    if( (uintptr_t)coefs & 15 ) return -2;
    if( b8x8 ) quant_kernel<64><<<(unsigned)( ( nblocks*8 + 255 ) / 256 ), 256, 0, st>>>( coefs, nblocks, mf, bias, nz );
    else       quant_kernel<16><<<(unsigned)( ( nblocks*2 + 255 ) / 256 ), 256, 0, st>>>( coefs, nblocks, mf, bias, nz );
    return 0;
}

// ---- frame_init_lowres_core with the reference's own signature (mc.c:484-507) ------------------------
// One thread = 4 output pixels of each of the four planes for LR_ROWS consecutive rows.  Aligned form: per input row two
// aligned words + the ninth byte (the next thread's first byte, an L1 hit); the third input row of one output row is
// the first of the next and stays in registers; vertical then horizontal byte averages on packed words (FILTER of
// mc.c:494 is avg(avg(a,b), avg(c,d)) with rounding byte averages), one 32-bit store per plane and row.
#ifndef LR_ROWS
#define LR_ROWS 4
#endif
template<bool ALIGNED>
__global__ void __launch_bounds__(128) lowres_core_kernel( const uint8_t *__restrict__ src, int ss, uint8_t *__restrict__ d0, uint8_t *__restrict__ dh,
                                                           uint8_t *__restrict__ dv, uint8_t *__restrict__ dc, int ds, int width, int height )
{
    const int x0 = ( blockIdx.x * blockDim.x + threadIdx.x ) * 4, y0 = blockIdx.y * LR_ROWS;
    if( x0 >= width ) return;
    if( ALIGNED && x0 + 4 <= width )
    {
        const uint8_t *r = src + (size_t)2*y0*ss + 2*x0;
        uint2 a = __ldg( (const uint2 *)r ); u32 a8 = __ldg( r + 8 );
#pragma unroll
        for( int j = 0; j < LR_ROWS; j++ )
        {
            if( y0 + j >= height ) break;
            const uint8_t *r1 = r + ss, *r2 = r1 + ss;
            const uint2 b = __ldg( (const uint2 *)r1 ), d = __ldg( (const uint2 *)r2 );
            const u32 b8 = __ldg( r1 + 8 ), d8 = __ldg( r2 + 8 );
            const u32 v0 = avg4( a.x, b.x ), v1 = avg4( a.y, b.y ), v8 = ( a8 + b8 + 1 ) >> 1;   // rows 2y, 2y+1
            const u32 w0 = avg4( b.x, d.x ), w1 = avg4( b.y, d.y ), w8 = ( b8 + d8 + 1 ) >> 1;   // rows 2y+1, 2y+2
            const u32 ve = __byte_perm( v0, v1, 0x6420 ), vo = __byte_perm( v0, v1, 0x7531 ), ve2 = ( ve >> 8 ) | ( v8 << 24 );
            const u32 we = __byte_perm( w0, w1, 0x6420 ), wo = __byte_perm( w0, w1, 0x7531 ), we2 = ( we >> 8 ) | ( w8 << 24 );
            const size_t o = (size_t)( y0 + j ) * ds + x0;
            *(u32 *)( d0 + o ) = avg4( ve, vo ); *(u32 *)( dh + o ) = avg4( vo, ve2 );
            *(u32 *)( dv + o ) = avg4( we, wo ); *(u32 *)( dc + o ) = avg4( wo, we2 );
            a = d; a8 = d8; r = r2;
        }
        return;
    }
#define FILT(a,b,c,d) ((((((a)+(b)+1)>>1)+(((c)+(d)+1)>>1)+1)>>1))
    for( int j = 0; j < LR_ROWS && y0 + j < height; j++ )
        for( int k = 0; k < 4 && x0 + k < width; k++ )
        {
            const int x = x0 + k, y = y0 + j;
            const uint8_t *r0 = src + (size_t)2*y*ss + 2*x, *r1 = r0 + ss, *r2 = r1 + ss;
            int a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2], c0 = r2[0], c1 = r2[1], c2 = r2[2];
            d0[(size_t)y*ds + x] = FILT( a0, b0, a1, b1 );
            dh[(size_t)y*ds + x] = FILT( a1, b1, a2, b2 );
            dv[(size_t)y*ds + x] = FILT( b0, c0, b1, c1 );
            dc[(size_t)y*ds + x] = FILT( b1, c1, b2, c2 );
        }
#undef FILT
}

// the same with 8 output pixels per thread (128-bit loads, 64-bit stores) for 16-byte aligned sources
__global__ void __launch_bounds__(128) lowres_core_kernel8( const uint8_t *__restrict__ src, int ss, uint8_t *__restrict__ d0, uint8_t *__restrict__ dh,
                                                            uint8_t *__restrict__ dv, uint8_t *__restrict__ dc, int ds, int width, int height )
{
    const int x0 = ( blockIdx.x * blockDim.x + threadIdx.x ) * 8, y0 = blockIdx.y * LR_ROWS;
    if( x0 >= width ) return;
    if( x0 + 8 <= width )
    {
        const uint8_t *r = src + (size_t)2*y0*ss + 2*x0;
        uint4 a = __ldg( (const uint4 *)r ); u32 a16 = __ldg( r + 16 );
#pragma unroll
        for( int j = 0; j < LR_ROWS; j++ )
        {
            if( y0 + j >= height ) break;
            const uint8_t *r1 = r + ss, *r2 = r1 + ss;
            const uint4 b = __ldg( (const uint4 *)r1 ), d = __ldg( (const uint4 *)r2 );
            const u32 b16 = __ldg( r1 + 16 ), d16 = __ldg( r2 + 16 );
            const size_t o = (size_t)( y0 + j ) * ds + x0;
#define LR_PAIR( P0, P1, DA, DB ) { \
            const u32 v0 = avg4( P0.x, P1.x ), v1 = avg4( P0.y, P1.y ), v2 = avg4( P0.z, P1.z ), v3 = avg4( P0.w, P1.w ); \
            const u32 v16 = ( P0##16 + P1##16 + 1 ) >> 1; \
            const u32 el = __byte_perm( v0, v1, 0x6420 ), eh = __byte_perm( v2, v3, 0x6420 ); \
            const u32 ol = __byte_perm( v0, v1, 0x7531 ), oh = __byte_perm( v2, v3, 0x7531 ); \
            const u32 e2l = __funnelshift_r( el, eh, 8 ), e2h = ( eh >> 8 ) | ( v16 << 24 ); \
            *(uint2 *)( DA + o ) = make_uint2( avg4( el, ol ), avg4( eh, oh ) ); \
            *(uint2 *)( DB + o ) = make_uint2( avg4( ol, e2l ), avg4( oh, e2h ) ); }
            LR_PAIR( a, b, d0, dh )
            LR_PAIR( b, d, dv, dc )
#undef LR_PAIR
            a = d; a16 = d16; r = r2;
        }
        return;
    }
#define FILT(a,b,c,d) ((((((a)+(b)+1)>>1)+(((c)+(d)+1)>>1)+1)>>1))
    for( int j = 0; j < LR_ROWS && y0 + j < height; j++ )
        for( int k = 0; k < 8 && x0 + k < width; k++ )
        {
            const int x = x0 + k, y = y0 + j;
            const uint8_t *r0 = src + (size_t)2*y*ss + 2*x, *r1 = r0 + ss, *r2 = r1 + ss;
            int a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2], c0 = r2[0], c1 = r2[1], c2 = r2[2];
            d0[(size_t)y*ds + x] = FILT( a0, b0, a1, b1 );
            dh[(size_t)y*ds + x] = FILT( a1, b1, a2, b2 );
            dv[(size_t)y*ds + x] = FILT( b0, c0, b1, c1 );
            dc[(size_t)y*ds + x] = FILT( b1, c1, b2, c2 );
        }
#undef FILT
}

int launch_lowres_core( const uint8_t *src, int sstride, uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc, int dstride,
                        int width, int height, cudaStream_t st )
{
    if( width <= 0 || height <= 0 ) return 0;
    const bool aligned16 = ( ( (uintptr_t)src | (uintptr_t)sstride ) & 15 ) == 0 &&
                           ( ( (uintptr_t)d0 | (uintptr_t)dh | (uintptr_t)dv | (uintptr_t)dc | (uintptr_t)dstride ) & 7 ) == 0;
    if( aligned16 )
    {
        dim3 grid8( ( ( width + 7 ) / 8 + 127 ) / 128, ( height + LR_ROWS - 1 ) / LR_ROWS );
        lowres_core_kernel8<<<grid8, 128, 0, st>>>( src, sstride, d0, dh, dv, dc, dstride, width, height );
        return 0;
    }
    dim3 grid( ( ( width + 3 ) / 4 + 127 ) / 128, ( height + LR_ROWS - 1 ) / LR_ROWS );
    const bool aligned = ( ( (uintptr_t)src | (uintptr_t)sstride ) & 7 ) == 0 &&
                         ( ( (uintptr_t)d0 | (uintptr_t)dh | (uintptr_t)dv | (uintptr_t)dc | (uintptr_t)dstride ) & 3 ) == 0;
    if( aligned ) lowres_core_kernel<true><<<grid, 128, 0, st>>>( src, sstride, d0, dh, dv, dc, dstride, width, height );
    else          lowres_core_kernel<false><<<grid, 128, 0, st>>>( src, sstride, d0, dh, dv, dc, dstride, width, height );
    return 0;
}
