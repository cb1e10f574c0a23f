// final.cu -- decision kernels of the lookahead cost engine (sm_100a): bidir candidates, intra/inter decision, AQ,
// packing and sums, i.e. the tail of slicetype_mb_cost (slicetype.c:616-652,703-712,758-790,946-989).
#include "la_block.cuh"

// ---- mode decision: the tail of slicetype_mb_cost for one block --------------------------------------
// What one block adds to the sums of its evaluation (slicetype.c:762-788); the caller accumulates.
struct BlockSums { int cost, cost_aq, row_aq, intra; };

// icost / invq: the block's intra cost and AQ factor, fetched by the caller ahead of the bidir candidates
template<class BE>
__device__ __forceinline__ BlockSums final_block( const BE &be, const FinalJob &J, const LaParams &P, int bx, int by, int blk, int icost, int invq )
{
    BlockSums S = { 0, 0, 0, 0 };
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
    if( !be.leader() ) return S;
    bcost += 4;                                         // lowres_penalty (slicetype.c:758)
    const bool score = !border || tiny;
    if( !J.b_bidir )
    {   // slicetype.c:762-773
        int b_intra = icost < bcost;
        if( b_intra ) { bcost = icost; list_used = 0; }
        if( score && b_intra ) S.intra = 1;
    }
    if( !J.b_intra_only )
    {   // slicetype.c:776-788
        int bcost_aq = P.aq ? ( bcost * invq + 128 ) >> 8 : bcost;
        S.row_aq = bcost_aq;
        if( score ) { S.cost = bcost; S.cost_aq = bcost_aq; }
    }
    J.lowres_costs[blk] = (uint16_t)( min( bcost, 0x3FFF ) + ( list_used << 14 ) );     // slicetype.c:790
    return S;
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
    const BlockSums S = final_block( be, J, P, bx, by, blk, J.b_bidir ? 0 : (int)J.intra[blk], P.aq && !J.b_intra_only ? (int)J.inv_qscale[blk] : 256 );
    if( be.g.r != 0 ) return;
    if( S.intra ) atomicAdd( &J.stats[2], 1 );
    if( !J.b_intra_only )
    {
        atomicAdd( &J.row_satds[by], S.row_aq );
        if( S.cost | S.cost_aq ) { atomicAdd( &J.stats[0], S.cost ); atomicAdd( &J.stats[1], S.cost_aq ); }
    }
}

// ---- thread-per-block decision kernel over CHUNKS of evaluations ---------------------------------------------------
// The B evaluations (p0, p1, b) of one batch that share the frame b AND the future reference p1 -- at bframes = 16 the
// 136 B evaluations of a lookahead step are 16 such families of 1..16 members -- also share the source block, the
// list-1 MV field and therefore the whole list-1 prediction get_ref( p1, mv(b -> p1) ), and their temporal-direct
// list-1 windows overlap (d1 = -mv(p1 -> p0) * (p1-b)/(p1-p0) is about the same motion for every p0).  The host sorts
// such families together (capi.cu group_finals) and cuts them into chunks of <= X264CU_F_CHUNK members; ONE thread
// handles one block for all members of a chunk: the source rows stay in registers, the list-1 prediction is built once
// (kept in registers, F_HOIST 1, or in shared memory, F_HOIST 2) and the list-1 windows of the later members hit in L1.
// Measured on B200 (profiles/r02/tuning.md): the per-evaluation kernel re-fetched every plane ~27x into L1 (ncu r01:
// 139 M L1 sectors per launch, hit rate 65 %, L1 data-pipe wavefronts 85 % of peak).
// Semantics per member are those of final_block (slicetype.c:616-652,703-712,758-790), bit for bit.
#ifndef X264CU_F_HOIST
#define X264CU_F_HOIST 2
#endif
// Which thread-per-block decision kernel launch_final() uses: 1 (default) = final_kernel_t, one CTA per (evaluation, tile),
// jobs in family order; 0 = final_kernel_c, one CTA per (chunk of a family, tile).  Measured on B200 (profiles/r02/tuning.md):
// the chunked kernel issues 27 % fewer L1 requests but keeps fewer warps resident (96 vs 80 registers) and runs in few long
// CTAs, so the L1 data pipe that bounds both kernels is 60 % busy instead of 86 %: 0.285 ms against 0.217 ms per step.
#ifndef X264CU_F_OLD
#define X264CU_F_OLD 1
#endif

template<bool SATD, int PITCH>
__global__ void __launch_bounds__(F_TILE_W*F_TILE_H, X264CU_MINB_FINAL) final_kernel_c( const FinalJob *__restrict__ jobs, const int2 *__restrict__ chunks, LaParams P )
{
#if X264CU_F_HOIST == 2
    __shared__ uint2 s_pred1[8][F_TILE_W*F_TILE_H];
#endif
    // the chunk's job descriptors, staged once per CTA: every field a member needs is then one shared-memory broadcast
    // away instead of a chain of dependent global loads (ncu r02a: 30 % of the stall samples sat on descriptor -> pointer
    // -> value chains in this latency-bound kernel)
    __shared__ __align__(16) FinalJob s_jobs[X264CU_F_CHUNK];
    const int2 ch = __ldg( chunks + blockIdx.y );           // first job, number of jobs (<= X264CU_F_CHUNK)
    {
        const unsigned long long *src = (const unsigned long long *)( jobs + ch.x );
        unsigned long long *dst = (unsigned long long *)s_jobs;
        for( int i = threadIdx.x; i < ch.y * (int)( sizeof(FinalJob) / 8 ); i += blockDim.x ) dst[i] = __ldg( src + i );
    }
    __syncthreads();
    int bx, by;
    if( !warp_tile_block( P.tiles_f, P.ntiles_f, P, bx, by ) ) return;
    const int blk = by * P.mbw + bx;
    typedef ThreadBE<SATD, false, PITCH, X264CU_FINAL_LD64, X264CU_FINAL_PTXPRED != 0> BE;
    BE be; be.P = &P;
    const int pitch = be.pitch(), pel = ( by*8 ) * pitch + bx*8;
    if( skipped_border( P, bx, by ) )
    {
        for( int z = 0; z < ch.y; z++ ) if( !s_jobs[z].b_intra_only ) s_jobs[z].lowres_costs[blk] = 0;
        return;
    }
    const bool border = bx == 0 || by == 0 || bx == P.mbw-1 || by == P.mbh-1;
    const bool score = !border || P.mbw <= 2 || P.mbh <= 2;
    // what every member of the chunk shares: the frame (source block, intra cost, AQ factor) and, for B evaluations, the
    // list-1 reference with its MV field
    const bool bidir = s_jobs[0].b_bidir != 0;
    const int icost = bidir ? 0 : (int)__ldg( s_jobs[0].intra + blk );
    const int invq = P.aq ? (int)__ldg( s_jobs[0].inv_qscale + blk ) : 256;
    int mv1x = 0, mv1y = 0, mc1 = X264CU_COST_MAX;
#if X264CU_F_HOIST == 1
    uint2 pred1[8];
#endif
    if( bidir )
    {
        const FinalJob &J0 = s_jobs[0];
        const uint8_t *src = J0.fenc + pel;
#pragma unroll
        for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * pitch ) );
        be.ref1 = J0.ref1 + pel;
        const int v = __ldg( (const int *)J0.mv[1] + blk );
        mv1x = (int16_t)( v & 0xffff ); mv1y = v >> 16;
        mc1 = __ldg( J0.mvcost[1] + blk );
#if X264CU_F_HOIST
        {   // the list-1 half of the L0+L1 candidate (slicetype.c:582-597): get_ref at mv1, or the half-pel plane for subme <= 1
            uint2 t[8];
            if( SATD )
            {
                QpelRows<false, PITCH, X264CU_FINAL_LD64, X264CU_FINAL_PTXPRED != 0> r1( be.ref1, P, mv1x, mv1y, 0, -1, 0 );
#pragma unroll
                for( int r = 0; r < 8; r++ ) t[r] = r1( r );
            }
            else
            {
                const int i1 = ( ( mv1x & 2 ) >> 1 ) + ( mv1y & 2 );
                const RowSrc<X264CU_FINAL_LD64> q1( be.ref1, ( mv1x >> 2 ) + ( mv1y >> 2 ) * pitch + i1 * (int)P.plane_size );
#pragma unroll
                for( int r = 0; r < 8; r++ ) t[r] = q1.row( r * ( pitch >> 2 ) );
            }
#pragma unroll
            for( int r = 0; r < 8; r++ )
            {
#if X264CU_F_HOIST == 1
                pred1[r] = t[r];
#else
                s_pred1[r][threadIdx.x] = t[r];
#endif
            }
        }
#endif
    }
    // this block's scalars of member z: list-0 MV and cost, and the MV the temporal-direct candidate is derived from.
    // They are fetched one member AHEAD, so that their latency hides behind the SATDs of the current member.
    auto fetch = [&]( int z, int &mv0, int &mc0, int &dmv )
    {
        const FinalJob &Jn = s_jobs[z];
        mv0 = 0; mc0 = X264CU_COST_MAX; dmv = 0;
        if( !Jn.b_intra_only ) { mv0 = __ldg( (const int *)Jn.mv[0] + blk ); mc0 = __ldg( Jn.mvcost[0] + blk ); }
        if( Jn.dmv_src ) dmv = __ldg( (const int *)Jn.dmv_src + blk );
    };
    int n_mv0, n_mc0, n_dmv;
    fetch( 0, n_mv0, n_mc0, n_dmv );
#pragma unroll 1
    for( int z = 0; z < ch.y; z++ )
    {
        const FinalJob &J = s_jobs[z];
        const int mv0 = n_mv0, mc0 = n_mc0, dmv = n_dmv;
        if( z + 1 < ch.y ) fetch( z + 1, n_mv0, n_mc0, n_dmv );
        int bcost = X264CU_COST_MAX, list_used = 0;
        if( !J.b_intra_only )
        {
            const int mv0x = (int16_t)( mv0 & 0xffff ), mv0y = mv0 >> 16;
            if( bidir )
            {
                be.ref0 = J.ref0 + pel;
                const int w = J.bipred_weight;
                int d0x = 0, d0y = 0, d1x = 0, d1y = 0;
                if( J.dmv_src )
                {   // temporal direct (slicetype.c:629-639)
                    MeGeo m;
                    me_limits( m, P, bx, by );
                    const int mrx = (int16_t)( dmv & 0xffff ), mry = dmv >> 16, dsf = J.dist_scale_factor;
                    d0x = ( mrx * dsf + 128 ) >> 8;
                    d0y = ( mry * dsf + 128 ) >> 8;
                    d1x = d0x - mrx; d1y = d0y - mry;
                    d0x = iclip3( d0x, m.sminx, m.smaxx ); d0y = iclip3( d0y, m.sminy, m.smaxy );
                    d1x = iclip3( d1x, m.sminx, m.smaxx ); d1y = iclip3( d1y, m.sminy, m.smaxy );
                    if( !SATD ) { d0x &= ~1; d0y &= ~1; d1x &= ~1; d1y &= ~1; }
                }
                // candidates in the reference's order, ties keep the earlier one: direct, zero (:644-652), L0, L1 (:708),
                // L0+L1 with the 5*lambda penalty (:711-712)
                {
                    const int cost = be.bidir( d0x, d0y, d1x, d1y, w );
                    if( cost < bcost ) { bcost = cost; list_used = 3; }
                }
                if( d0x | d0y | d1x | d1y )
                {
                    const int cz = be.bidir_zero( w );
                    if( cz < bcost ) { bcost = cz; list_used = 3; }
                }
                if( mc0 < bcost ) { bcost = mc0; list_used = 1; }
                if( mc1 < bcost ) { bcost = mc1; list_used = 2; }
                if( mv0x | mv0y | mv1x | mv1y )
                {
                    int cost;
#if X264CU_F_HOIST
#if X264CU_F_HOIST == 1
#define PRED1( r ) pred1[r]
#else
#define PRED1( r ) s_pred1[r][threadIdx.x]
#endif
                    if( SATD )
                    {
                        QpelRows<false, PITCH, X264CU_FINAL_LD64, X264CU_FINAL_PTXPRED != 0> r0( be.ref0, P, mv0x, mv0y, 0, -1, 0 );
                        cost = satd8x8_t( be.fenc, [&]( int r ) { const uint2 a = r0( r ), b = PRED1( r ); return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
                    }
                    else
                    {
                        const int i0 = ( ( mv0x & 2 ) >> 1 ) + ( mv0y & 2 );
                        const RowSrc<X264CU_FINAL_LD64> q0( be.ref0, ( mv0x >> 2 ) + ( mv0y >> 2 ) * pitch + i0 * (int)P.plane_size );
                        cost = sad8x8_t( be.fenc, [&]( int r ) { const uint2 a = q0.row( r * ( pitch >> 2 ) ), b = PRED1( r ); return make_uint2( biavg4( a.x, b.x, w ), biavg4( a.y, b.y, w ) ); } );
                    }
#undef PRED1
#else
                    cost = be.bidir( mv0x, mv0y, mv1x, mv1y, w );
#endif
                    cost += 5;
                    if( cost < bcost ) { bcost = cost; list_used = 3; }
                }
            }
            else if( mc0 < bcost ) { bcost = mc0; list_used = 1; }
        }
        bcost += 4;                                         // lowres_penalty (slicetype.c:758)
        if( !bidir )
        {   // slicetype.c:762-773
            const int b_intra = icost < bcost;
            if( b_intra ) { bcost = icost; list_used = 0; }
            if( score && b_intra ) atomicAdd( &J.stats[2], 1 );
        }
        if( !J.b_intra_only )
        {   // slicetype.c:776-788
            const int bcost_aq = P.aq ? ( bcost * invq + 128 ) >> 8 : bcost;
            atomicAdd( &J.row_satds[by], bcost_aq );
            if( score )
            {
                atomicAdd( &J.stats[0], bcost );
                atomicAdd( &J.stats[1], bcost_aq );
            }
        }
        J.lowres_costs[blk] = (uint16_t)( min( bcost, 0x3FFF ) + ( list_used << 14 ) );     // slicetype.c:790
    }
}

#if X264CU_F_OLD
// one CTA per (evaluation, block tile): many short CTAs keep the L1 data pipe, which bounds the decision kernels, busiest
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
    // every lane of a warp that owns a tile stays until the sums are accumulated: the 32 blocks' contributions are added up
    // with warp reductions and leave as one atomic per sum and block row (integer adds: any order gives the same result) --
    // per-thread atomics were 3.3 M same-address L2 operations per 1080p launch (154 evaluations x 8160 blocks x 3)
    const int wt = blockIdx.x * ( blockDim.x >> 5 ) + ( threadIdx.x >> 5 ), lane = threadIdx.x & 31;
    if( wt >= P.ntiles_f ) return;
    const uint32_t td = __ldg( P.tiles_f + wt );
    const int lw = td >> 24;
    const int bx = (int)( td & 0xfff ) + ( lane & ( ( 1 << lw ) - 1 ) ), by = (int)( ( td >> 12 ) & 0xfff ) + ( lane >> lw );
    const bool inside = bx < P.mbw && by < P.mbh;
    const int blk = by * P.mbw + bx;
    ThreadBE<SATD, false, PITCH, X264CU_FINAL_LD64, X264CU_FINAL_PTXPRED != 0> be; be.P = &P;
    const int pel = ( by*8 ) * be.pitch() + bx*8;
#pragma unroll 1
    for( int z = blockIdx.y * F_JOBS; z < min( njobs, ( blockIdx.y + 1 ) * F_JOBS ); z++ )
    {
        const FinalJob J = jobs[z];
        BlockSums S = { 0, 0, 0, 0 };
        if( inside )
        {
            if( skipped_border( P, bx, by ) ) { if( !J.b_intra_only ) J.lowres_costs[blk] = 0; }
            else
            {
                // the two scalars the decision needs last are requested first: their latency hides behind the candidates
                const int icost = J.b_bidir ? 0 : (int)__ldg( J.intra + blk );
                const int invq = P.aq && !J.b_intra_only ? (int)__ldg( J.inv_qscale + blk ) : 256;
                const uint8_t *src = J.fenc + pel;
                if( !J.b_intra_only )
                {
#pragma unroll
                    for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * be.pitch() ) );
                }
                be.ref0 = J.ref0 + pel; be.ref1 = J.ref1 + pel; be.ref_w = be.ref0;
                S = final_block( be, J, P, bx, by, blk, icost, invq );
            }
        }
        if( !J.b_bidir )
        {
            const int n = __reduce_add_sync( 0xffffffffu, S.intra );
            if( lane == 0 && n ) atomicAdd( &J.stats[2], n );
        }
        if( !J.b_intra_only )
        {
            const int c = __reduce_add_sync( 0xffffffffu, S.cost ), ca = __reduce_add_sync( 0xffffffffu, S.cost_aq );
            if( lane == 0 && ( c | ca ) ) { atomicAdd( &J.stats[0], c ); atomicAdd( &J.stats[1], ca ); }
            for( int j = 0; j < ( 32 >> lw ); j++ )
            {   // one block row of the warp tile at a time (two for the 16-wide tiles)
                const int r = __reduce_add_sync( 0xffffffffu, ( lane >> lw ) == j ? S.row_aq : 0 );
                if( lane == ( j << lw ) && r ) atomicAdd( &J.row_satds[by], r );
            }
        }
    }
}

#endif

void launch_final( const FinalJob *d_jobs, int njobs, const int2 *d_chunks, int nchunks, const LaParams &P, cudaStream_t st )
{
    if( use_thread_mapping( P, njobs ) )
    {
        const int warps_per_cta = F_TILE_W * F_TILE_H / 32;
        dim3 grid( ( P.ntiles_f + warps_per_cta - 1 ) / warps_per_cta, nchunks );
#if X264CU_F_OLD
        grid.y = njobs;
#define FINAL_T( PITCH ) { if( P.b_satd ) final_kernel_t<true, PITCH><<<grid, F_TILE_W*F_TILE_H, 0, st>>>( d_jobs, njobs, P ); \
                           else           final_kernel_t<false, PITCH><<<grid, F_TILE_W*F_TILE_H, 0, st>>>( d_jobs, njobs, P ); }
#else
#define FINAL_T( PITCH ) { if( P.b_satd ) final_kernel_c<true, PITCH><<<grid, F_TILE_W*F_TILE_H, 0, st>>>( d_jobs, d_chunks, P ); \
                           else           final_kernel_c<false, PITCH><<<grid, F_TILE_W*F_TILE_H, 0, st>>>( d_jobs, d_chunks, P ); }
#endif
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
