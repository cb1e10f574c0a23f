// mbtree.cu -- MB-tree propagation on the device (first "next" row).
//
//   mbtree_propagate_kernel   mbtree_propagate_cost (common/mc.c:511-525) fused with both
//                             mbtree_propagate_list calls (mc.c:527-598) of one
//                             macroblock_tree_propagate step (encoder/slicetype.c:1051-1089)
//   mbtree_finish_kernel      macroblock_tree_finish (slicetype.c:1029-1049)
//
// The reference accumulates into uint16 with a saturating add (MC_CLIP_ADD, mc.h:29).  All addends
// are non-negative, so min(min(a+x,M)+y,M) == min(a+x+y,M): we accumulate in 32-bit with atomics (order
// free, deterministic) and clamp to 32767 whenever the value is read.  Float arithmetic uses explicit
// round-to-nearest intrinsics in the reference's operation order so that no FMA contraction changes a
// rounding (the reference C is compiled without FMA).
#include "lookahead.cuh"

#define LOWRES_COST_MASK 0x3FFF
#define LOWRES_COST_SHIFT 14

__device__ __forceinline__ void clip_add( u32 *dst, int v ) { if( v ) atomicAdd( dst, (u32)v ); }

__global__ void __launch_bounds__(256) mbtree_propagate_kernel( MbtreeJob J, LaParams P )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= P.nmb ) return;
    const int bx = i % P.mbw, by = i / P.mbw;
    // mbtree_propagate_cost
    const int intra_cost = J.intra[i];
    const int lc = J.lowres_costs[i];
    const int inter_cost = min( intra_cost, lc & LOWRES_COST_MASK );
    const int prop_in = J.referenced ? (int)min( J.prop_b[i], 32767u ) : 0;
    const float propagate_intra = __int2float_rn( intra_cost * (int)J.inv_qscale[i] );
    const float propagate_amount = __fadd_rn( __int2float_rn( prop_in ), __fmul_rn( propagate_intra, J.fps_factor ) );
    const float propagate_num = __int2float_rn( intra_cost - inter_cost );
    const float propagate_denom = __int2float_rn( intra_cost );
    const int amount = min( __float2int_rz( __fadd_rn( __fdiv_rn( __fmul_rn( propagate_amount, propagate_num ), propagate_denom ), 0.5f ) ), 32767 );
    // mbtree_propagate_list for list 0 and (B frames) list 1
    const int lists_used = lc >> LOWRES_COST_SHIFT;
#pragma unroll
    for( int list = 0; list < 2; list++ )
    {
        if( !J.mvs[list] || !( lists_used & ( 1 << list ) ) ) continue;
        u32 *ref_costs = J.prop_ref[list];
        int listamount = amount;
        if( lists_used == 3 ) listamount = ( listamount * J.bipred_weight[list] + 32 ) >> 6;
        const int v = __ldg( (const int *)J.mvs[list] + i );
        int x = (int16_t)( v & 0xffff ), y = v >> 16;
        if( !v ) { clip_add( &ref_costs[i], listamount ); continue; }
        const unsigned mbx = (unsigned)( ( x >> 5 ) + bx ), mby = (unsigned)( ( y >> 5 ) + by );
        const unsigned idx0 = mbx + mby * P.mbw, idx2 = idx0 + P.mbw;
        x &= 31; y &= 31;
        const int w0 = ( ( 32-y ) * ( 32-x ) * listamount + 512 ) >> 10, w1 = ( ( 32-y ) * x * listamount + 512 ) >> 10;
        const int w2 = ( y * ( 32-x ) * listamount + 512 ) >> 10, w3 = ( y * x * listamount + 512 ) >> 10;
        const unsigned width = P.mbw, height = P.mbh;
        if( mby < height )
        {
            if( mbx < width ) clip_add( &ref_costs[idx0], w0 );
            if( mbx + 1 < width ) clip_add( &ref_costs[idx0+1], w1 );
        }
        if( mby + 1 < height )
        {
            if( mbx < width ) clip_add( &ref_costs[idx2], w2 );
            if( mbx + 1 < width ) clip_add( &ref_costs[idx2+1], w3 );
        }
    }
}

void launch_mbtree_propagate( const MbtreeJob &J, const LaParams &P, cudaStream_t st )
{
    mbtree_propagate_kernel<<<( P.nmb + 255 ) / 256, 256, 0, st>>>( J, P );
}

// x264_log2 (common/base.h:226-230) with the tables of common/tables.c:66-90
__device__ __forceinline__ float x264_log2_dev( u32 x, const float *lut )
{
    const int lz = __clz( x );
    return __fadd_rn( lut[( x << lz >> 24 ) & 0x7f], __int2float_rn( 31 - lz ) );
}

__global__ void __launch_bounds__(256) mbtree_finish_kernel( const uint16_t *__restrict__ intra, const uint16_t *__restrict__ inv_qscale,
        const u32 *__restrict__ prop, const float *__restrict__ qp_offset_aq, float *__restrict__ qp_offset,
        const float *__restrict__ log2_lut, int fps_factor, float weightdelta, float strength, int nmb )
{
    __shared__ float s_lut[128];
    if( threadIdx.x < 128 ) s_lut[threadIdx.x] = log2_lut[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= nmb ) return;
    const int intra_cost = ( (int)intra[i] * (int)inv_qscale[i] + 128 ) >> 8;
    if( intra_cost )
    {
        const int propagate_cost = ( (int)min( prop[i], 32767u ) * fps_factor + 128 ) >> 8;
        const float log2_ratio = __fadd_rn( __fsub_rn( x264_log2_dev( intra_cost + propagate_cost, s_lut ), x264_log2_dev( intra_cost, s_lut ) ), weightdelta );
        qp_offset[i] = __fsub_rn( qp_offset_aq[i], __fmul_rn( strength, log2_ratio ) );
    }
}

void launch_mbtree_finish( const uint16_t *intra, const uint16_t *inv_qscale, const u32 *prop, const float *qp_offset_aq, float *qp_offset,
                           const float *log2_lut, int fps_factor, float weightdelta, float strength, int nmb, cudaStream_t st )
{
    mbtree_finish_kernel<<<( nmb + 255 ) / 256, 256, 0, st>>>( intra, inv_qscale, prop, qp_offset_aq, qp_offset, log2_lut, fps_factor, weightdelta, strength, nmb );
}
