// aq_weights.cu -- the two producers right before the lookahead cost path, sm_100a.
//
//   aq_kernel               x264_adaptive_quant_frame for aq-mode 0..3 (ratecontrol.c:226-415): per macroblock the AC
//                           energy of luma 16x16 + both chroma 8x8 (ac_energy_mb :262-294, pixel_var pixel.c:158-175),
//                           qp_adj = strength * (x264_log2(energy) - 14.427) (:388-399), i_inv_qscale_factor =
//                           x264_exp2fix8(qp_adj) (base.h:216-224), and the frame's i_pixel_sum / i_pixel_ssd
//   weights_guess_kernel    the float prologue of x264_weights_analyse(h, fenc, ref, 1) (slicetype.c:284-360,399-420)
//   weight_cost_kernel      weight_cost_luma without and with the candidate weight (slicetype.c:191-222), optionally on
//                           the motion-compensated reference (weight_cost_init_luma :76-103)
//   weights_decide_kernel   the decision epilogue (slicetype.c:170-189,434-466)
//   weight_plane_dev_kernel x264_weight_scale_plane (frame.c:825-840) under a weight that lives in device memory
//
// Everything that x264 computes in float is computed with explicit round-to-nearest single operations in the same
// order (no FMA contraction), so the results equal the reference's bit for bit (tests/test_gpu_aq_weights.py).
// The streaming kernel (aq_kernel) is HBM-bound: 3.1 MB read per 1080p frame, 49 KB written.
#include "la_block.cuh"

// ------------------------------------------------------------------------------------------------
// adaptive quantisation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float x264_log2_dev( u32 x, const float *lut )
{   // base.h:226-230; x264_log2_lz_lut[lz] == 31 - lz
    const int lz = __clz( x );
    return __fadd_rn( lut[( x << lz >> 24 ) & 0x7f], (float)( 31 - lz ) );
}
__device__ __forceinline__ int x264_exp2fix8_dev( float x, const uint8_t *lut )
{   // base.h:216-224: int i = x*(-64.f/6.f) + 512.5f
    const int i = __float2int_rz( __fadd_rn( __fmul_rn( x, -64.f / 6.f ), 512.5f ) );
    if( i < 0 ) return 0;
    if( i > 1023 ) return 0xffff;
    return ( ( lut[i & 63] + 256 ) << ( i >> 6 ) ) >> 8;
}

// powf( x, 0.125f ) of ratecontrol.c:361 for x = energy + 1 (an integer below 2^24, exact in float), bit-identical with the
// host libm the reference links: the candidate sqrt(sqrt(sqrt((double)x))) rounded to float (three correctly rounded
// IEEE square roots: the same bits on host and device) differs from glibc's powf -- which is accurate to ~0.5 ulp but not
// correctly rounded -- on ~0.07 % of the inputs; capi.cu finds exactly those inputs by running the host's powf over the
// whole domain once and hands them over as a sorted exception table (pow_x: float bits of x, pow_v: powf's result).
struct PowTab { const u32 *x; const float *v; int n; };
__device__ __forceinline__ float pow_eighth_dev( u32 energy, const PowTab &t )
{
    const float x = __fadd_rn( __fmul_rn( __uint2float_rn( energy ), 1.f ), 1.f );      // energy * bit_depth_correction + 1
    const u32 key = __float_as_uint( x );
    int lo = 0, hi = t.n;
    while( lo < hi ) { const int mid = ( lo + hi ) >> 1; if( __ldg( t.x + mid ) < key ) lo = mid + 1; else hi = mid; }
    if( lo < t.n && __ldg( t.x + lo ) == key ) return __ldg( t.v + lo );
    return __double2float_rn( sqrt( sqrt( sqrt( (double)x ) ) ) );
}

// One warp per macroblock: lanes 0-15 own the 16 luma rows, lanes 16-23 the 8 rows of interleaved NV12 chroma.
// Pixels beyond the picture replicate the last row / column / chroma pair (x264_frame_expand_border_mod16).
__global__ void __launch_bounds__(256) aq_kernel( const uint8_t *__restrict__ luma, int ls, const uint8_t *__restrict__ chroma, int cs,
                                                  int W, int H, int mbw, int nmb, int mode, float strength,
                                                  const float *__restrict__ quant_offsets, const float *__restrict__ log2_lut,
                                                  const uint8_t *__restrict__ exp2_lut,
                                                  float *__restrict__ qp_offset, uint16_t *__restrict__ inv_qscale,
                                                  unsigned long long *__restrict__ pixstats, PowTab pw )
{
    __shared__ unsigned long long s_acc[6];
    __shared__ float s_lut[128];
    __shared__ uint8_t s_exp[64];
    if( threadIdx.x < 6 ) s_acc[threadIdx.x] = 0;
    if( threadIdx.x < 128 ) s_lut[threadIdx.x] = log2_lut[threadIdx.x];
    if( threadIdx.x < 64 ) s_exp[threadIdx.x] = exp2_lut[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, mb = blockIdx.x * ( blockDim.x >> 5 ) + ( threadIdx.x >> 5 );
    if( mb < nmb )
    {
        const int mx = mb % mbw, my = mb / mbw;
        u32 v[6] = {0, 0, 0, 0, 0, 0};      // sum_y, ssd_y, sum_u, ssd_u, sum_v, ssd_v
        if( lane < 16 )
        {
            const int y = min( 16*my + lane, H-1 ), x0 = 16*mx;
            const uint8_t *row = luma + (size_t)y * ls;
            u32 w[4];
            if( x0 + 16 <= W && ( ( (uintptr_t)luma | (uintptr_t)ls ) & 15 ) == 0 )
            {
                const uint4 q = __ldg( (const uint4 *)( row + x0 ) );
                w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            }
            else
            {
#pragma unroll
                for( int k = 0; k < 4; k++ )
                {
                    w[k] = 0;
#pragma unroll
                    for( int b = 0; b < 4; b++ ) w[k] |= (u32)row[min( x0 + 4*k + b, W-1 )] << ( 8*b );
                }
            }
#pragma unroll
            for( int k = 0; k < 4; k++ ) { v[0] = __dp4a( w[k], 0x01010101u, v[0] ); v[1] = __dp4a( w[k], w[k], v[1] ); }
        }
        else if( lane < 24 )
        {
            const int cw = W >> 1, ch = H >> 1;                      // chroma size in U/V pairs
            const int y = min( 8*my + lane - 16, ch-1 ), x0 = 8*mx;
            const uint8_t *row = chroma + (size_t)y * cs;
            u32 w[4];
            if( x0 + 8 <= cw && ( ( (uintptr_t)chroma | (uintptr_t)cs ) & 15 ) == 0 )
            {
                const uint4 q = __ldg( (const uint4 *)( row + 2*x0 ) );
                w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            }
            else
            {
#pragma unroll
                for( int k = 0; k < 4; k++ )
                {
                    w[k] = 0;
#pragma unroll
                    for( int b = 0; b < 4; b++ ) w[k] |= (u32)row[2*min( x0 + 2*k + ( b >> 1 ), cw-1 ) + ( b & 1 )] << ( 8*b );
                }
            }
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                const u32 u = w[k] & 0x00ff00ffu, vv = ( w[k] >> 8 ) & 0x00ff00ffu;
                v[2] = __dp4a( u, 0x01010101u, v[2] );  v[3] = __dp4a( u, u, v[3] );
                v[4] = __dp4a( vv, 0x01010101u, v[4] ); v[5] = __dp4a( vv, vv, v[5] );
            }
        }
#pragma unroll
        for( int k = 0; k < 6; k++ )
#pragma unroll
            for( int o = 16; o > 0; o >>= 1 ) v[k] += __shfl_xor_sync( 0xffffffffu, v[k], o );
        if( lane == 0 )
        {
            // ac_energy_var (ratecontrol.c:226-236): ssd - (sum*sum >> shift), all uint32 after the 64-bit product
            u32 energy = v[1] - (u32)( ( (unsigned long long)v[0] * v[0] ) >> 8 );
            energy += v[3] - (u32)( ( (unsigned long long)v[2] * v[2] ) >> 6 );
            energy += v[5] - (u32)( ( (unsigned long long)v[4] * v[4] ) >> 6 );
            if( mode >= 2 && strength != 0.f )
                qp_offset[mb] = pow_eighth_dev( energy, pw );        // first pass of the auto-variance modes (ratecontrol.c:355-366)
            else if( mode >= 1 )
            {
                float qp_adj;
                if( strength != 0.f )
                {   // ratecontrol.c:388-399
                    qp_adj = __fmul_rn( strength, __fsub_rn( x264_log2_dev( max( energy, 1u ), s_lut ), 14.427f ) );
                    if( quant_offsets ) qp_adj = __fadd_rn( qp_adj, quant_offsets[mb] );
                }
                else
                    qp_adj = quant_offsets ? quant_offsets[mb] : 0.f;          // :318-333
                qp_offset[mb] = qp_adj;
                inv_qscale[mb] = ( strength != 0.f || quant_offsets ) ? (uint16_t)x264_exp2fix8_dev( qp_adj, s_exp ) : (uint16_t)256;
            }
            else
                qp_offset[mb] = 0.f;
#pragma unroll
            for( int k = 0; k < 6; k++ ) atomicAdd( &s_acc[k], (unsigned long long)v[k] );
        }
    }
    __syncthreads();
    if( threadIdx.x < 6 && s_acc[threadIdx.x] ) atomicAdd( &pixstats[threadIdx.x], s_acc[threadIdx.x] );
}

// aq-mode 2 / 3, second and third step (ratecontrol.c:363-398).  The frame averages are float sums accumulated in raster
// order by the reference, so ONE thread adds the nmb values in that order (8160 dependent adds at 1080p, ~25 us; the
// auto-variance modes are not x264's default and this keeps them bit-exact), then every macroblock gets its offset.
__global__ void aq_auto_reduce_kernel( const float *__restrict__ qp, int nmb, float strength, float *__restrict__ out3 )
{
    if( threadIdx.x | blockIdx.x ) return;
    float a = 0.f, p = 0.f;
    int i = 0;
    for( ; i + 4 <= nmb && ( ( (uintptr_t)qp & 15 ) == 0 ); i += 4 )
    {
        const float4 q = __ldg( (const float4 *)( qp + i ) );
        a = __fadd_rn( a, q.x ); p = __fadd_rn( p, __fmul_rn( q.x, q.x ) );
        a = __fadd_rn( a, q.y ); p = __fadd_rn( p, __fmul_rn( q.y, q.y ) );
        a = __fadd_rn( a, q.z ); p = __fadd_rn( p, __fmul_rn( q.z, q.z ) );
        a = __fadd_rn( a, q.w ); p = __fadd_rn( p, __fmul_rn( q.w, q.w ) );
    }
    for( ; i < nmb; i++ ) { const float q = qp[i]; a = __fadd_rn( a, q ); p = __fadd_rn( p, __fmul_rn( q, q ) ); }
    a = __fdiv_rn( a, (float)nmb );
    p = __fdiv_rn( p, (float)nmb );
    out3[0] = __fmul_rn( strength, a );                                                              // strength = f_aq_strength * avg_adj
    out3[1] = __fsub_rn( a, __fdiv_rn( __fmul_rn( 0.5f, __fsub_rn( p, 14.f ) ), a ) );               // avg_adj - 0.5f * (avg_adj_pow2 - 14.f) / avg_adj
    out3[2] = strength;                                                                              // bias_strength
}

__global__ void __launch_bounds__(256) aq_auto_apply_kernel( int mode, int nmb, const float *__restrict__ red3, const float *__restrict__ quant_offsets,
                                                             const uint8_t *__restrict__ exp2_lut, float *__restrict__ qp_offset, uint16_t *__restrict__ inv_qscale )
{
    __shared__ uint8_t s_exp[64];
    if( threadIdx.x < 64 ) s_exp[threadIdx.x] = exp2_lut[threadIdx.x];
    __syncthreads();
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if( mb >= nmb ) return;
    const float strength = red3[0], avg_adj = red3[1], bias_strength = red3[2];
    float q = qp_offset[mb];
    float r = __fmul_rn( strength, __fsub_rn( q, avg_adj ) );
    if( mode == 3 )      // + bias_strength * (1.f - 14.f / (qp_adj * qp_adj))
        r = __fadd_rn( r, __fmul_rn( bias_strength, __fsub_rn( 1.f, __fdiv_rn( 14.f, __fmul_rn( q, q ) ) ) ) );
    if( quant_offsets ) r = __fadd_rn( r, quant_offsets[mb] );
    qp_offset[mb] = r;
    inv_qscale[mb] = (uint16_t)x264_exp2fix8_dev( r, s_exp );
}

// strength: f_aq_strength * 1.0397f for mode 1 (ratecontrol.c:371), f_aq_strength itself for modes 2 / 3
int launch_aq( const uint8_t *luma, int ls, const uint8_t *chroma, int cs, int W, int H, const LaParams &P, int mode, float strength,
               const float *quant_offsets, const float *log2_lut, const uint8_t *exp2_lut, float *qp_offset, uint16_t *inv_qscale,
               unsigned long long *pixstats, const u32 *pow_x, const float *pow_v, int pow_n, float *red3, cudaStream_t st )
{
    PowTab pw = { pow_x, pow_v, pow_n };
    aq_kernel<<<( P.nmb + 7 ) / 8, 256, 0, st>>>( luma, ls, chroma, cs, W, H, P.mbw, P.nmb, mode, strength, quant_offsets, log2_lut, exp2_lut,
                                                  qp_offset, inv_qscale, pixstats, pw );
    if( mode < 2 || strength == 0.f ) return 1;
    aq_auto_reduce_kernel<<<1, 32, 0, st>>>( qp_offset, P.nmb, strength, red3 );
    aq_auto_apply_kernel<<<( P.nmb + 255 ) / 256, 256, 0, st>>>( mode, P.nmb, red3, quant_offsets, exp2_lut, qp_offset, inv_qscale );
    return 3;
}

// ------------------------------------------------------------------------------------------------
// weighted-prediction analysis
// ------------------------------------------------------------------------------------------------
// wres layout (int32[8] per (fenc, dist)): [0] enabled [1] scale [2] denom [3] offset  <- the result
//                                          [4] state: 0 no candidate (early termination), 1 candidate in [1..3]
//                                          [5] origscore [6] weighted score (sums of weight_cost_kernel) [7] minscore
__device__ __forceinline__ void final_pixstats( const unsigned long long *raw, int mbw, int mbh, unsigned long long &sum, unsigned long long &ssd )
{   // ratecontrol.c:406-414 for the luma plane
    const unsigned long long wh = (unsigned long long)( 16*mbw ) * ( 16*mbh );
    sum = raw[0];
    ssd = raw[1] - ( sum * sum + wh / 2 ) / wh;
}

__global__ void weights_guess_kernel( const unsigned long long *__restrict__ st_fenc, const unsigned long long *__restrict__ st_ref,
                                      int mbw, int mbh, int *__restrict__ wres )
{
    if( threadIdx.x | blockIdx.x ) return;
    unsigned long long fsum, fssd, rsum, rssd;
    final_pixstats( st_fenc, mbw, mbh, fsum, fssd );
    final_pixstats( st_ref, mbw, mbh, rsum, rssd );
    wres[0] = 0; wres[1] = 1; wres[2] = 0; wres[3] = 0; wres[4] = 0; wres[5] = 0; wres[6] = 0; wres[7] = 0;
    // slicetype.c:299-307
    const int zero_bias = !rssd;
    const float fenc_var = __ull2float_rn( fssd + zero_bias ), ref_var = __ull2float_rn( rssd + zero_bias );
    const float guess_scale = __fsqrt_rn( __fdiv_rn( fenc_var, ref_var ) );
    const float npix = (float)( ( 16*mbh ) * ( 16*mbw ) );
    const float fenc_mean = __fdiv_rn( __fdiv_rn( __ull2float_rn( fsum + zero_bias ), npix ), 1.f );
    const float ref_mean  = __fdiv_rn( __fdiv_rn( __ull2float_rn( rsum + zero_bias ), npix ), 1.f );
    // early termination (:335-339)
    if( fabsf( __fsub_rn( ref_mean, fenc_mean ) ) < 0.5f && fabsf( __fsub_rn( 1.f, guess_scale ) ) < 1.f/128.f ) return;
    // weight_get_h264( round( guess_scale * 128 ), 0, w ) (:62-74); round() is the double-precision C function
    int denom = 7, scale = (int)round( (double)__fmul_rn( guess_scale, 128.f ) );
    while( denom > 0 && scale > 127 ) { denom--; scale >>= 1; }
    scale = min( scale, 127 );
    // the single candidate of the lookahead (:399-420)
    int cur_scale = scale;
    int cur_offset = __float2int_rz( __fadd_rn( __fsub_rn( fenc_mean, __fdiv_rn( __fmul_rn( ref_mean, (float)cur_scale ), (float)( 1 << denom ) ) ), 0.5f ) );
    if( cur_offset < -128 || cur_offset > 127 )
    {
        cur_offset = iclip3( cur_offset, -128, 127 );
        float v = __fadd_rn( __fdiv_rn( __fmul_rn( (float)( 1 << denom ), __fsub_rn( fenc_mean, (float)cur_offset ) ), ref_mean ), 0.5f );
        v = v < 0.f ? 0.f : v > 127.f ? 127.f : v;        // x264_clip3f
        cur_scale = __float2int_rz( v );
    }
    wres[1] = cur_scale; wres[2] = denom; wres[3] = iclip3( cur_offset, -128, 127 ); wres[4] = 1;
    wres[7] = scale;       // the un-clamped starting scale: minscale before any candidate wins (:356-358)
}

// weight_cost_luma for both hypotheses in one pass over the blocks, one thread per 8x8 block
template<bool SATD, int PITCH>
__global__ void __launch_bounds__(128) weight_cost_kernel( const uint8_t *__restrict__ fencF, const uint8_t *__restrict__ refF,
                                                           const int16_t *__restrict__ mvs, const uint16_t *__restrict__ intra,
                                                           int *__restrict__ wres, LaParams P )
{
    if( !wres[4] ) return;                                   // early termination: no candidate to score
    const int bx = blockIdx.x * 32 + ( threadIdx.x & 31 ), by = blockIdx.y * 4 + ( threadIdx.x >> 5 );
    u32 c0 = 0, c1 = 0;
    if( bx < P.mbw && by < P.mbh )
    {
        const int blk = by * P.mbw + bx;
        ThreadBE<SATD, true, PITCH, false, false> be; be.P = &P;
        const int pel = ( by*8 ) * be.pitch() + bx*8;
        const uint8_t *src = fencF + pel;
#pragma unroll
        for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * be.pitch() ) );
        be.ref0 = refF + pel; be.ref1 = NULL; be.ref_w = be.ref0;
        int mvx = 0, mvy = 0;
        if( mvs ) { const int v = __ldg( (const int *)mvs + blk ); mvx = (int16_t)( v & 0xffff ); mvy = v >> 16; }
        MeGeo m;
        m.wdenom = -1; m.wscale = 0; m.woffset = 0;
        const int cmp0 = SATD ? be.qsatd( m, mvx, mvy ) : be.qsad1( m, mvx, mvy );
        m.wscale = wres[1]; m.wdenom = wres[2]; m.woffset = wres[3];
        const int cmp1 = SATD ? be.qsatd( m, mvx, mvy ) : be.qsad1( m, mvx, mvy );
        const int ic = intra[blk];
        c0 = min( cmp0, ic ); c1 = min( cmp1, ic );
    }
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 ) { c0 += __shfl_xor_sync( 0xffffffffu, c0, o ); c1 += __shfl_xor_sync( 0xffffffffu, c1, o ); }
    if( ( threadIdx.x & 31 ) == 0 ) { atomicAdd( (u32 *)&wres[5], c0 ); atomicAdd( (u32 *)&wres[6], c1 ); }
}

__device__ __forceinline__ int ue_size_tab_dev( int i ) { return i ? 2 * ( 31 - __clz( i ) ) + 1 : 1; }      // bitstream.h:201-219
__device__ __forceinline__ int bs_size_se_dev( int v )
{   // bitstream.h:291-299
    int tmp = 1 - v*2;
    if( tmp < 0 ) tmp = v*2;
    return tmp < 256 ? ue_size_tab_dev( tmp ) : ue_size_tab_dev( tmp >> 8 ) + 16;
}

__global__ void weights_decide_kernel( int *__restrict__ wres, int numslices )
{
    if( threadIdx.x | blockIdx.x ) return;
    if( !wres[4] ) return;
    const u32 origscore = (u32)wres[5];
    u32 minscore = origscore;
    int minscale = wres[7], mindenom = wres[2], minoff = 0, found = 0;
    wres[0] = 0;
    if( !origscore ) { wres[1] = 1; wres[2] = 0; wres[3] = 0; wres[7] = 0; return; }                         // slicetype.c:387-388
    // weight_slice_header_cost (:170-189), lambda 1, luma
    const u32 s = (u32)wres[6] + numslices * ( 10 + ue_size_tab_dev( mindenom + 1 ) * 2 + 2 * ( bs_size_se_dev( wres[1] ) + bs_size_se_dev( wres[3] ) ) );
    if( s < minscore ) { minscore = s; minscale = wres[1]; minoff = wres[3]; found = 1; }                    // COPY4_IF_LT (:432)
    while( mindenom > 0 && !( minscale & 1 ) ) { mindenom--; minscale >>= 1; }                               // :444-451
    const bool off = !found || ( minscale == 1 << mindenom && minoff == 0 ) || __fdiv_rn( (float)minscore, (float)origscore ) > 0.998f;
    wres[0] = !off; wres[1] = off ? 1 : minscale; wres[2] = off ? 0 : mindenom; wres[3] = off ? 0 : minoff;
    wres[7] = (int)minscore;
}

void launch_weights_analyse( const uint8_t *fencF, const uint8_t *refF, const int16_t *mvs, const uint16_t *intra,
                             const unsigned long long *st_fenc, const unsigned long long *st_ref, int *wres, int numslices,
                             const LaParams &P, cudaStream_t st )
{
    weights_guess_kernel<<<1, 32, 0, st>>>( st_fenc, st_ref, P.mbw, P.mbh, wres );
    dim3 grid( ( P.mbw + 31 ) / 32, ( P.mbh + 3 ) / 4 );
#define WCOST( PITCH ) { if( P.b_satd ) weight_cost_kernel<true, PITCH><<<grid, 128, 0, st>>>( fencF, refF, mvs, intra, wres, P ); \
                         else           weight_cost_kernel<false, PITCH><<<grid, 128, 0, st>>>( fencF, refF, mvs, intra, wres, P ); }
    if( P.stride == X264CU_PITCH_A )      WCOST( X264CU_PITCH_A )
    else if( P.stride == X264CU_PITCH_B ) WCOST( X264CU_PITCH_B )
    else                        WCOST( 0 )
#undef WCOST
    weights_decide_kernel<<<1, 32, 0, st>>>( wres, numslices );
}

// x264_weight_scale_plane (frame.c:825-840) of all of plane F incl. its border, under a device-resident weight
__global__ void weight_plane_dev_kernel( const u32 *__restrict__ src, u32 *__restrict__ dst, int64_t nwords, const int *__restrict__ wres )
{
    if( !wres[0] ) return;
    const int scale = wres[1], denom = wres[2], offset = wres[3];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( i < nwords ) dst[i] = weight4( src[i], scale, denom, offset );
}

void launch_weight_plane_dev( const uint8_t *src, uint8_t *dst, int64_t bytes, const int *wres, cudaStream_t st )
{
    const int64_t nwords = bytes >> 2;
    weight_plane_dev_kernel<<<(unsigned)( ( nwords + 255 ) / 256 ), 256, 0, st>>>( (const u32 *)src, (u32 *)dst, nwords, wres );
}
