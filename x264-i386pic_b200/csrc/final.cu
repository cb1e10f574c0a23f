// final.cu -- decision kernels of the lookahead cost engine (sm_100a): bidir candidates, intra/inter decision, AQ,
// packing and sums, i.e. the tail of slicetype_mb_cost (slicetype.c:616-652,703-712,758-790,946-989).
#include "la_block.cuh"

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

// F_JOBS consecutive evaluations of the job table per CTA, same block tile.  The batch lists the evaluations of one p0
// back to back, so consecutive jobs read the same two reference frames and the tile's windows could stay in L1 from
// one job to the next -- measured on B200 it does not pay (F_JOBS 4 / 8 / 16: 0.247 / 0.289 / 0.266 ms against 0.225 for 1:
// fewer, longer CTAs balance worse over the SMs), so the default is one evaluation per CTA.
#ifndef F_JOBS
#define F_JOBS 1
#endif
template<bool SATD, int PITCH>
__global__ void __launch_bounds__(F_TILE_W*F_TILE_H, X264CU_MINB_FINAL) final_kernel_t( const FinalJob *__restrict__ jobs, int njobs, LaParams P )
{
    int bx, by;
    if( !warp_tile_block( P.tiles_f, P.ntiles_f, P, bx, by ) ) return;
    const int blk = by * P.mbw + bx;
    ThreadBE<SATD, false, PITCH, X264CU_FINAL_LD64 != 0, X264CU_FINAL_PTXPRED != 0> be; be.P = &P;
    const int pel = ( by*8 ) * be.pitch() + bx*8;
#pragma unroll 1
    for( int z = blockIdx.y * F_JOBS; z < min( njobs, ( blockIdx.y + 1 ) * F_JOBS ); z++ )
    {
        const FinalJob J = jobs[z];
        if( skipped_border( P, bx, by ) ) { if( !J.b_intra_only ) J.lowres_costs[blk] = 0; continue; }
        const uint8_t *src = J.fenc + pel;
        if( !J.b_intra_only )
        {
#pragma unroll
            for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * be.pitch() ) );
        }
        be.ref0 = J.ref0 + pel; be.ref1 = J.ref1 + pel; be.ref_w = be.ref0;
        final_block( be, J, P, bx, by, blk );
    }
}

void launch_final( const FinalJob *d_jobs, int njobs, const LaParams &P, cudaStream_t st )
{
    if( use_thread_mapping( P, njobs ) )
    {
        const int warps_per_cta = F_TILE_W * F_TILE_H / 32;
        dim3 grid( ( P.ntiles_f + warps_per_cta - 1 ) / warps_per_cta, ( njobs + F_JOBS - 1 ) / F_JOBS );
#define FINAL_T( PITCH ) { if( P.b_satd ) final_kernel_t<true, PITCH><<<grid, F_TILE_W*F_TILE_H, 0, st>>>( d_jobs, njobs, P ); \
                           else           final_kernel_t<false, PITCH><<<grid, F_TILE_W*F_TILE_H, 0, st>>>( d_jobs, njobs, P ); }
        if( P.stride == X264CU_PITCH_A )      FINAL_T( X264CU_PITCH_A )
        else if( P.stride == X264CU_PITCH_B ) FINAL_T( X264CU_PITCH_B )
        else                        FINAL_T( 0 )
#undef FINAL_T
    }
    else
    {
        dim3 grid( ( P.nmb + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, njobs );
        final_kernel_g<<<grid, CTA_THREADS, 0, st>>>( d_jobs, P );
    }
}
