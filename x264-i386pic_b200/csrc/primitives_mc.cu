// primitives_mc.cu -- the checkasm surfaces of the path's pixel fetch / prediction code as batched entry points (sm_100a):
//   mc_luma / get_ref     common/mc.c:198-249 (tools/checkasm.c:1234-1307): qpel fetch from four phase planes + weight
//   pixel_avg (12 sizes)  common/mc.c:49-103 (checkasm.c:1325-1353): bipred average with any weight -63..127
//   mc_weight             common/mc.c:117-137 (checkasm.c:1355-1429): scale/denom/offset incl. denom == 0
//   intra_{sad,satd}_x3_8x8c  common/pixel.c:542-556, predict.c:221-304 (checkasm.c:754)
//   quant_4x4x4           common/quant.c:75-86
// The lookahead kernels carry their own in-register forms of the same arithmetic (la_block.cuh QpelRows / biavg4 /
// weight4, lookahead.cu intra_kernel); these entry points expose it at the reference's function granularity so that it
// can be compared function by function.  One thread per output pixel quad / per block: streaming, HBM-bound.
#include "primitives.cuh"

static __constant__ uint8_t c_mc_ref0[16] = {0,1,1,1,0,1,1,1,2,3,3,3,0,1,1,1};   // x264_hpel_ref0, tables.c:183
static __constant__ uint8_t c_mc_ref1[16] = {0,0,1,0,2,2,3,2,2,2,3,2,2,2,3,2};   // x264_hpel_ref1, tables.c:184

__device__ __forceinline__ int clip_px( int v ) { return min( max( v, 0 ), 255 ); }
__device__ __forceinline__ int weight1( int v, int scale, int denom, int offset )
{   // mc.c:117-118 opscale / opscale_noden
    return clip_px( denom >= 1 ? ( ( v*scale + ( 1 << ( denom-1 ) ) ) >> denom ) + offset : v*scale + offset );
}

// ---- mc_luma: out[i] = w x h block i, packed ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mc_luma_kernel( const uint8_t *__restrict__ buf, int64_t po0, int64_t po1, int64_t po2, int64_t po3, int stride,
                                                       int64_t n, const int64_t *__restrict__ org, const int16_t *__restrict__ mv, int w, int h,
                                                       int wscale, int wdenom, int woffset, uint8_t *__restrict__ out )
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, per = (int64_t)w * h;
    if( t >= n * per ) return;
    const int64_t i = t / per;
    const int x = (int)( t % per ) % w, y = (int)( t % per ) / w;
    const int mvx = mv[2*i], mvy = mv[2*i+1];
    const int qidx = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
    const int64_t po[4] = { po0, po1, po2, po3 };
    const int64_t off = org[i] + (int64_t)( ( mvy >> 2 ) + y ) * stride + ( mvx >> 2 ) + x;
    int v = buf[po[c_mc_ref0[qidx]] + off + ( ( mvy & 3 ) == 3 ? stride : 0 )];
    if( qidx & 5 )                                            // mc.c:206-213 / :232-236: average of two half-pel planes
        v = ( v + buf[po[c_mc_ref1[qidx]] + off + ( ( mvx & 3 ) == 3 )] + 1 ) >> 1;
    if( wdenom >= 0 ) v = weight1( v, wscale, wdenom, woffset );
    out[t] = (uint8_t)v;
}

int launch_mc_luma( const uint8_t *buf, const int64_t plane_off[4], int stride, int64_t n, const int64_t *org, const int16_t *mv,
                    int w, int h, int wscale, int wdenom, int woffset, uint8_t *out, cudaStream_t st )
{
    if( w < 1 || h < 1 || w > 64 || h > 64 ) return -1;
    if( n <= 0 ) return 0;
    const int64_t total = n * w * h;
    mc_luma_kernel<<<(unsigned)( ( total + 255 ) / 256 ), 256, 0, st>>>( buf, plane_off[0], plane_off[1], plane_off[2], plane_off[3], stride, n, org, mv,
                                                                         w, h, wscale, wdenom, woffset, out );
    return 0;
}

// ---- pixel_avg: (a*w + b*(64-w) + 32) >> 6 clipped, or (a+b+1)>>1 for w == 32 (mc.c:49-87) ------------------------------
__global__ void __launch_bounds__(256) pixel_avg_kernel( const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2, int64_t n,
                                                         const int64_t *__restrict__ off1, const int64_t *__restrict__ off2, int w, int h, int weight,
                                                         uint8_t *__restrict__ out )
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, per = (int64_t)w * h;
    if( t >= n * per ) return;
    const int64_t i = t / per;
    const int x = (int)( t % per ) % w, y = (int)( t % per ) / w;
    const int a = p1[off1[i] + (int64_t)y * s1 + x], b = p2[off2[i] + (int64_t)y * s2 + x];
    out[t] = (uint8_t)( weight == 32 ? ( a + b + 1 ) >> 1 : clip_px( ( a * weight + b * ( 64 - weight ) + 32 ) >> 6 ) );
}

int launch_pixel_avg( const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n, const int64_t *off1, const int64_t *off2,
                      int w, int h, int weight, uint8_t *out, cudaStream_t st )
{
    if( w < 1 || h < 1 || w > 64 || h > 64 ) return -1;
    if( n <= 0 ) return 0;
    const int64_t total = n * w * h;
    pixel_avg_kernel<<<(unsigned)( ( total + 255 ) / 256 ), 256, 0, st>>>( p1, s1, p2, s2, n, off1, off2, w, h, weight, out );
    return 0;
}

// ---- mc_weight over a plane (mc.c:119-137; x264_weight_scale_plane frame.c:825-840 calls it in 16-line strips) ----------
__global__ void __launch_bounds__(256) mc_weight_kernel( const uint8_t *__restrict__ src, int sstride, uint8_t *__restrict__ dst, int dstride,
                                                         int width, int height, int scale, int denom, int offset )
{
    const int x4 = ( blockIdx.x * blockDim.x + threadIdx.x ) * 4, y = blockIdx.y;
    if( x4 >= width || y >= height ) return;
#pragma unroll
    for( int k = 0; k < 4; k++ )
        if( x4 + k < width ) dst[(int64_t)y * dstride + x4 + k] = (uint8_t)weight1( src[(int64_t)y * sstride + x4 + k], scale, denom, offset );
}

int launch_mc_weight( const uint8_t *src, int sstride, uint8_t *dst, int dstride, int width, int height, int scale, int denom, int offset, cudaStream_t st )
{
    if( width <= 0 || height <= 0 ) return 0;
    if( denom < 0 ) return -1;
    dim3 grid( ( ( width + 3 ) / 4 + 255 ) / 256, height );
    mc_weight_kernel<<<grid, 256, 0, st>>>( src, sstride, dst, dstride, width, height, scale, denom, offset );
    return 0;
}

// ---- intra_{sad,satd}_x3_8x8c: res = { DC, H, V } (pixel.c:552-553) --------------------------------------------------------
// fenc block at off1 of plane 1; its neighbours (top row, left column) are read around off2 of plane 2, as x264_predict_8x8c_*
// read them around fdec (predict.c:221-304).  One thread per block; SATD = two satd_8x4 halves (pixel.c:290-330).
__device__ __forceinline__ int satd8x8_pred( const uint2 fenc[8], const uint2 pred[8] )
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
            diff_row8( fenc[4*h+i], pred[4*h+i], d );
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
__device__ __forceinline__ int sad8x8_pred( const uint2 fenc[8], const uint2 pred[8] )
{
    u32 s = 0;
#pragma unroll
    for( int r = 0; r < 8; r++ ) s += sad8( fenc[r], pred[r] );
    return (int)s;
}

__global__ void __launch_bounds__(128) intra_x3_8x8c_kernel( int b_satd, const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2,
                                                             int64_t n, const int64_t *__restrict__ off1, const int64_t *__restrict__ off2, int32_t *__restrict__ res )
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n ) return;
    const uint8_t *f = p1 + off1[i], *d = p2 + off2[i];
    uint2 fenc[8], pred[8];
    u32 left[8];
#pragma unroll
    for( int r = 0; r < 8; r++ ) { fenc[r] = load8u( f + (int64_t)r * s1 ); left[r] = d[(int64_t)r * s2 - 1]; }
    const uint2 top = load8u( d - s2 );
    // DC (predict.c:221-253): per 4x4 quadrant from the 4-pixel sums of the top row halves and the left column halves
    const u32 s0 = __vsadu4( top.x, 0 ), s1_ = __vsadu4( top.y, 0 ), s2_ = left[0] + left[1] + left[2] + left[3], s3 = left[4] + left[5] + left[6] + left[7];
    const u32 dc0 = ( s0 + s2_ + 4 ) >> 3, dc1 = ( s1_ + 2 ) >> 2, dc2 = ( s3 + 2 ) >> 2, dc3 = ( s1_ + s3 + 4 ) >> 3;
#pragma unroll
    for( int r = 0; r < 8; r++ ) pred[r] = r < 4 ? make_uint2( dc0 * 0x01010101u, dc1 * 0x01010101u ) : make_uint2( dc2 * 0x01010101u, dc3 * 0x01010101u );
    res[3*i+0] = b_satd ? satd8x8_pred( fenc, pred ) : sad8x8_pred( fenc, pred );
#pragma unroll
    for( int r = 0; r < 8; r++ ) pred[r] = make_uint2( left[r] * 0x01010101u, left[r] * 0x01010101u );      // H (predict.c:256-266)
    res[3*i+1] = b_satd ? satd8x8_pred( fenc, pred ) : sad8x8_pred( fenc, pred );
#pragma unroll
    for( int r = 0; r < 8; r++ ) pred[r] = top;                                                              // V (predict.c:267-277)
    res[3*i+2] = b_satd ? satd8x8_pred( fenc, pred ) : sad8x8_pred( fenc, pred );
}

int launch_intra_x3_8x8c( int b_satd, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n, const int64_t *off1, const int64_t *off2,
                          int32_t *res, cudaStream_t st )
{
    if( n <= 0 ) return 0;
    intra_x3_8x8c_kernel<<<(unsigned)( ( n + 127 ) / 128 ), 128, 0, st>>>( b_satd, p1, s1, p2, s2, n, off1, off2, res );
    return 0;
}

// ---- quant_4x4x4 (quant.c:75-86): four 4x4 blocks, bit j of the result = block j has a non-zero coefficient ----------------
__global__ void __launch_bounds__(128) quant_4x4x4_kernel( int16_t *__restrict__ coefs, int64_t ngroups, const uint16_t *__restrict__ mf,
                                                           const uint16_t *__restrict__ bias, int32_t *__restrict__ nza )
{
    __shared__ uint16_t s_mf[16], s_bias[16];
    if( threadIdx.x < 16 ) { s_mf[threadIdx.x] = mf[threadIdx.x]; s_bias[threadIdx.x] = bias[threadIdx.x]; }
    __syncthreads();
    // 4 threads per group: thread j quantises block j; the group's mask is assembled with shuffles
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < ngroups * 4;
    int nz = 0;
    if( live )
    {
        int16_t *c = coefs + t * 16;
        __align__(16) int16_t v[16];
        ( (uint4 *)v )[0] = ( (const uint4 *)c )[0]; ( (uint4 *)v )[1] = ( (const uint4 *)c )[1];
#pragma unroll
        for( int k = 0; k < 16; k++ )
        {   // QUANT_ONE (quant.c:50-57): u32 arithmetic, sign-symmetric
            const int x = v[k];
            const int q = x > 0 ? (int)( ( (u32)s_bias[k] + (u32)x ) * s_mf[k] >> 16 ) : -(int)( ( (u32)s_bias[k] + (u32)( -x ) ) * s_mf[k] >> 16 );
            v[k] = (int16_t)q;
            nz |= q;
        }
        ( (uint4 *)c )[0] = ( (uint4 *)v )[0]; ( (uint4 *)c )[1] = ( (uint4 *)v )[1];
    }
    int bit = ( nz != 0 ) << ( threadIdx.x & 3 );
    bit |= __shfl_xor_sync( 0xffffffffu, bit, 1 );
    bit |= __shfl_xor_sync( 0xffffffffu, bit, 2 );
    if( live && !( threadIdx.x & 3 ) ) nza[t >> 2] = bit;
}

int launch_quant_4x4x4( int16_t *coefs, int64_t ngroups, const uint16_t *mf, const uint16_t *bias, int32_t *nza, cudaStream_t st )
{
    if( ngroups <= 0 ) return 0;
    if( (uintptr_t)coefs & 15 ) return -2;
    quant_4x4x4_kernel<<<(unsigned)( ( ngroups * 4 + 127 ) / 128 ), 128, 0, st>>>( coefs, ngroups, mf, bias, nza );
    return 0;
}
