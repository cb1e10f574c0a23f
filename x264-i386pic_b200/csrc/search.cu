// search.cu -- motion search kernels of the lookahead cost engine (sm_100a).
//
//   search_kernel_t / _g   x264_me_search_ref for PIXEL_8x8 (me.c:182-420,774-797,865-992) under the MV-predictor
//                          rule of slicetype.c:658-697, one pass of the dependency-free schedule
//   search_kernel_wave     the same block search in the reference's raster-serial order, one anti-diagonal per launch
#include "la_block.cuh"

#define FPEL(v) (((v)+2)>>2)
// half-pel diamond around an odd centre with pairwise shared fetches (1, la_block.cuh hpel4_pairs) or as four get_ref candidates (0)
// a dedicated instantiation of the thread-per-block search for pass 0 (1) or one kernel for every pass (0)
// pass 0: the SATD stage keeps the centre's rows and fetches one window per quarter-pel neighbour (1) or runs the generic get_ref (0)
#ifndef X264CU_EVEN_SATD
#define X264CU_EVEN_SATD 1
#endif
#ifndef X264CU_FIRST_PASS
#define X264CU_FIRST_PASS 1
#endif
#ifndef X264CU_HPEL_PAIRS
#define X264CU_HPEL_PAIRS 1
#endif

// refine_subpel (me.c:865-992), p_halfpel_thresh == NULL, b_refine_qpel == 0, no chroma
// EVEN: every centre is known to have even components (pass 0: the start is full-pel and the half-pel diamond moves by 2)
template<class BE, bool EVEN = false>
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
#if X264CU_HPEL_EVEN
            if constexpr( BE::HPEL_EVEN )
            {
#if X264CU_HPEL_PAIRS
                // both forms give the same result for an even, unweighted centre, so the vote only decides what the warp runs:
                // pass 1 (zero predictor) has even centres only and takes the cheaper single-plane form as a whole
                const bool even = !( ( omx | omy ) & 1 ) && m.wdenom < 0;
                if( __all_sync( __activemask(), even ) ) be.hpel4_even( omx, omy, tab, m, bcost, best );
                else be.hpel4_pairs( omx, omy, tab, m, bcost, best );
#else
                if( !( ( omx | omy ) & 1 ) && m.wdenom < 0 ) be.hpel4_even( omx, omy, tab, m, bcost, best );
                else be.qsad_best( 4, [&]( int k, int &x, int &y ) { x = omx + ( k == 2 ? -2 : k == 3 ? 2 : 0 ); y = omy + ( k == 0 ? -2 : k == 1 ? 2 : 0 ); },
                                   tab, m, bcost, best );
#endif
            }
            else
#endif
            be.qsad_best( 4, [&]( int k, int &x, int &y ) { x = omx + ( k == 2 ? -2 : k == 3 ? 2 : 0 ); y = omy + ( k == 0 ? -2 : k == 1 ? 2 : 0 ); },
                          tab, m, bcost, best );
            if( best < 0 ) break;
            bmx = omx + ( best == 2 ? -2 : best == 3 ? 2 : 0 );
            bmy = omy + ( best == 0 ? -2 : best == 1 ? 2 : 0 );
        }
    }
    if constexpr( BE::PAR4 && BE::SATD5 )
    {   // eight slots: the re-measurement and the quarter-pel diamond are one round
        const bool diamond = qpel_iters && !( bmy <= m.sminy || bmy >= m.smaxy || bmx <= m.sminx || bmx >= m.smaxx );
        if( be.satd() )
        {
            const int omx = bmx, omy = bmy;
            int best = -1;
            be.qsatd_best5( omx, omy, diamond, tab, m, bcost, best );
            if( best >= 0 ) { bmx = omx + ( best == 2 ? -1 : best == 3 ? 1 : 0 ); bmy = omy + ( best == 0 ? -1 : best == 1 ? 1 : 0 ); }
        }
    }
    else if constexpr( BE::PAR4 )
    {
        if( be.satd() )   // mbcmp != fpelcmp: re-measure the winner with SATD (me.c:925-929)
            bcost = be.qsatd( m, bmx, bmy ) + mvcost( tab, m, bmx, bmy );
        if( qpel_iters && !( bmy <= m.sminy || bmy >= m.smaxy || bmx <= m.sminx || bmx >= m.smaxx ) )
        {   // the four quarter-pel neighbours in one round (me.c:947-963, one iteration at la subme 4)
            const int omx = bmx, omy = bmy;
            int best = -1;
            be.qsatd_best4( [&]( int k, int &x, int &y ) { x = omx + ( k == 2 ? -1 : k == 3 ? 1 : 0 ); y = omy + ( k == 0 ? -1 : k == 1 ? 1 : 0 ); },
                            0, tab, m, bcost, best );
            if( best >= 0 ) { bmx = omx + ( best == 2 ? -1 : best == 3 ? 1 : 0 ); bmy = omy + ( best == 0 ? -1 : best == 1 ? 1 : 0 ); }
        }
    }
    else
    {   // the SATD re-measurement of the winner (me.c:925-929, k = -1) and the quarter-pel diamond (me.c:947-963, k = 0..3;
        // the lookahead runs one iteration, so no direction is excluded) share ONE call site: the SATD body is ~600
        // instructions and the kernel is instruction-cache sensitive (ncu: 16 % no-instruction stalls)
        const int omx = bmx, omy = bmy;
        const bool diamond = qpel_iters && !( bmy <= m.sminy || bmy >= m.smaxy || bmx <= m.sminx || bmx >= m.smaxx );
        auto generic = [&]()
        {
#pragma unroll 1
            for( int k = be.satd() ? -1 : 0; k < ( diamond ? 4 : 0 ); k++ )
            {
                const int mx = omx + ( k == 2 ? -1 : k == 3 ? 1 : 0 ), my = omy + ( k == 0 ? -1 : k == 1 ? 1 : 0 );
                const int cost = be.qsatd( m, mx, my ) + mvcost( tab, m, mx, my );
                if( k < 0 ) bcost = cost;
                else if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; }
            }
        };
#if X264CU_EVEN_SATD
        if constexpr( BE::HPEL_EVEN )
        {   // an even centre is ONE plane's rows, and each of its four quarter-pel neighbours is the average of exactly those rows
            // with one other window (mc.c:221-249 over the four even phases): the centre is fetched once and kept
            bool all_even = EVEN;
#if X264CU_EVEN_SATD == 2
            // later passes: a warp whose centres all happen to be even (and unweighted) takes the same path
            if constexpr( !EVEN ) all_even = be.satd() && __all_sync( __activemask(), !( ( omx | omy ) & 1 ) && m.wdenom < 0 );
#endif
            if( all_even )
            {
                if( be.satd() )
                {
                    uint2 K[8];
                    be.centre_rows( omx, omy, K );
#pragma unroll 1
                    for( int k = -1; k < ( diamond ? 4 : 0 ); k++ )
                    {
                        const int mx = omx + ( k == 2 ? -1 : k == 3 ? 1 : 0 ), my = omy + ( k == 0 ? -1 : k == 1 ? 1 : 0 );
                        const int cost = be.qsatd_even( K, omx, omy, mx, my ) + mvcost( tab, m, mx, my );
                        if( k < 0 ) bcost = cost;
                        else if( cost < bcost ) { bcost = cost; bmx = mx; bmy = my; }
                    }
                }
            }
            else generic();
        }
        else
#endif
        generic();
    }
}

// x264_me_search_ref (me.c:182-420,774-797) for one block; result mv (qpel) and cost
template<class BE, bool FIRST = false>
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
        int pmv_cost;
        if constexpr( BE::PAR4 )
        {   // the predictor and its candidates in one round of four (two when all four neighbours survive the filter)
            int cx[5], cy[5], nc = 1;
            cx[0] = bpred_mx; cy[0] = bpred_my;
#pragma unroll
            for( int i = 0; i < 4; i++ )
            {
                int mx = mvc[i][0], my = mvc[i][1];
                if( i >= i_mvc || ( !mx && !my ) || ( mx == pmvx && my == pmvy ) ) continue;
                cx[nc] = iclip3( mx, 4*m.fminx, 4*m.fmaxx ); cy[nc] = iclip3( my, 4*m.fminy, 4*m.fmaxy ); nc++;
            }
            int best = 0;
            be.qsad_best0( nc, [&]( int k, int &x, int &y ) { x = cx[k]; y = cy[k]; }, tab, m, bpred_cost, best, pmv_cost );
            bpred_mx = cx[best]; bpred_my = cy[best];
        }
        else
        {
        // the predictor itself (i = -1, me.c:219-222) and its candidates (me.c:230-246) through one call site
        pmv_cost = 0;
#pragma unroll 1
        for( int i = -1; i < i_mvc; i++ )
        {   // x264_predictor_clip (common.h:790-805): drop zero and == pmv, clip the rest
            int mx = i < 0 ? pmvx : mvc[i][0], my = i < 0 ? pmvy : mvc[i][1];
            if( i >= 0 )
            {
                if( ( !mx && !my ) || ( mx == pmvx && my == pmvy ) ) continue;
                mx = iclip3( mx, 4*m.fminx, 4*m.fmaxx ); my = iclip3( my, 4*m.fminy, 4*m.fmaxy );
            }
            const int cost = be.qsad1( m, mx, my ) + mvcost( tab, m, mx, my );
            if( i < 0 ) pmv_cost = cost;
            if( cost < bpred_cost ) { bpred_cost = cost; bpred_mx = mx; bpred_my = my; }
        }
        }
        bmx = FPEL( bpred_mx ); bmy = FPEL( bpred_my );
        // me.c:258-274 with the two possible fullpel probes folded into one call site
        const bool probe_start = ( bpred_mx & 3 ) | ( bpred_my & 3 );
        if( !probe_start ) bcost = bpred_cost;
        const bool probe_zero = ( pmvx | pmvy ) && ( bmx | bmy );
        const int sx = bmx, sy = bmy;
        if constexpr( BE::PAR4 )
        {
            const int np = (int)probe_start + (int)probe_zero;
            if( np )
            {
                int best = -1;
                be.fpel_best( np, [&]( int k, int &x, int &y ) { const bool z = k || !probe_start; x = z ? 0 : sx; y = z ? 0 : sy; }, tab, m, bcost, best );
                if( best >= 0 ) { const bool z = best || !probe_start; bmx = z ? 0 : sx; bmy = z ? 0 : sy; }
            }
        }
        else
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
            bmx += hex2x( dir+1 ); bmy += hex2y( dir+1 );
            for( int i = ( P.me_range >> 1 ) - 1; i > 0 && in_range( m, bmx, bmy ); i-- )
            {
                const int ox = bmx, oy = bmy, base = dir;
                best = -1;
                be.fpel_best( 3, [&]( int k, int &x, int &y ) { x = ox + hex2x( base+k ); y = oy + hex2y( base+k ); }, tab, m, bcost, best );
                if( best < 0 ) break;
                dir += best - 1;
                dir = dir < 0 ? dir + 6 : dir >= 6 ? dir - 6 : dir;      // mod6m1 (me.c:53)
                bmx = ox + hex2x( dir+1 ); bmy = oy + hex2y( dir+1 );
            }
        }
        // square refine (me.c:404-418)
        {
            const int ox = bmx, oy = bmy;
            int bs = -1;
            be.square8( ox, oy, tab, m, bcost, bs );
            bmx = ox + sq1x( bs+1 ); bmy = oy + sq1y( bs+1 );
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
        refine_subpel<BE, FIRST>( be, m, P, tab, mx, my, cost, 1, 1 );      // subpel_iterations[4] = {0,2,1,1}
    }
    out_mx = mx; out_my = my; out_cost = cost;
}

// one block of one search pass (slicetype.c:654-702)
// FIRST: the pass is known at compile time to have no predictor field (pass 0 of the dependency-free schedule): the predictor
// is zero and the candidate list empty for every block, so the candidate loop, the probes and the odd-centre half-pel path fold away
template<class BE, bool FIRST = false>
__device__ __forceinline__ int search_block( const BE &be, const SearchJob &J, const LaParams &P, const CostTab &tab,
                                             int bx, int by, int blk, int right_mv = 0 )
{
    MeGeo m;
    m.wscale = J.wscale; m.wdenom = J.wdenom; m.woffset = J.woffset;
    me_limits( m, P, bx, by );
    // MV predictor from the previous pass's field (slicetype.c:658-680): right, below, below-left, below-right
    int mvc[4][2] = {{0,0},{0,0},{0,0},{0,0}}, i_mvc = 0;
    if( !FIRST && J.mv_in )
    {
        const int *f = (const int *)J.mv_in + blk;
#define MVC_SET( v ) { int v_ = (v); mvc[i_mvc][0] = (int16_t)( v_ & 0xffff ); mvc[i_mvc][1] = v_ >> 16; i_mvc++; }
        if constexpr( BE::PAR4 )
        {   // serial kernel: the right neighbour is this warp's previous block; the row below is written by another
            // warp of the same launch and each MV word is its own "ready" flag (see search_kernel_serial)
            if( bx < P.mbw-1 ) MVC_SET( right_mv )
            if( by < P.mbh-1 )
            {
                int v0, v1, v2;
                be.wait_mv3( f + P.mbw, 1 | ( bx > 0 ? 2 : 0 ) | ( bx < P.mbw-1 ? 4 : 0 ), v0, v1, v2 );
                MVC_SET( v0 )
                if( bx > 0 ) MVC_SET( v1 )
                if( bx < P.mbw-1 ) MVC_SET( v2 )
            }
        }
        else
        {
            if( bx < P.mbw-1 ) MVC_SET( __ldg( f + 1 ) )
            if( by < P.mbh-1 )
            {
                MVC_SET( __ldg( f + P.mbw ) )
                if( bx > 0 ) MVC_SET( __ldg( f + P.mbw-1 ) )
                if( bx < P.mbw-1 ) MVC_SET( __ldg( f + P.mbw+1 ) )
            }
        }
#undef MVC_SET
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
        me_search<BE, FIRST>( be, m, P, tab, mvc, i_mvc, mx, my, cost );
        cost -= tab[0];                                // slicetype.c:695
        if( mx | my ) cost += 5;                       // 5 * lambda (:696-697)
    }
    const int packed = ( mx & 0xffff ) | ( my << 16 );
    if( be.leader() )
    {
        J.cost_out[blk] = cost;
        if constexpr( BE::PAR4 ) __stcg( (int *)J.mv_out + blk, packed );      // straight to L2, where the row above polls it
        else ( (int *)J.mv_out )[blk] = packed;
    }
    return packed;
}

// a job whose weight lives in device memory (x264cu_weights_analyse ran on the same stream, no host round trip)
__device__ __forceinline__ void resolve_device_weight( SearchJob &J )
{
    if( !J.wdev ) return;
    if( __ldg( J.wdev ) ) { J.wscale = __ldg( J.wdev + 1 ); J.wdenom = __ldg( J.wdev + 2 ); J.woffset = __ldg( J.wdev + 3 ); }
    else                  { J.wdenom = -1; J.ref_w = J.ref; }
}

__global__ void __launch_bounds__(CTA_THREADS) search_kernel_g( const SearchJob *__restrict__ jobs, LaParams P )
{
    const CostTab tab = stage_cost_tab( P );
    SearchJob J = jobs[blockIdx.y];
    resolve_device_weight( J );
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

template<bool SATD, bool WEIGHTED, int PITCH, bool FIRST = false>
__global__ void __launch_bounds__(S_TILE_W*S_TILE_H, X264CU_MINB_SEARCH) search_kernel_t( const SearchJob *__restrict__ jobs, LaParams P )
{
    const CostTab tab = stage_cost_tab( P );
    SearchJob J = jobs[blockIdx.y];
    resolve_device_weight( J );
    int bx, by;
    if( !warp_tile_block( P.tiles_s, P.ntiles_s, P, bx, by ) ) return;
    const int blk = by * P.mbw + bx;
    if( skipped_border( P, bx, by ) ) { ( (int *)J.mv_out )[blk] = 0; J.cost_out[blk] = 0; return; }
    ThreadBE<SATD, WEIGHTED, PITCH, X264CU_SEARCH_LD64, X264CU_SEARCH_PTXPRED != 0> be; be.P = &P;
    const int pel = ( by*8 ) * be.pitch() + bx*8;
    const uint8_t *src = J.fenc + pel;
#pragma unroll
    for( int r = 0; r < 8; r++ ) be.fenc[r] = __ldg( (const uint2 *)( src + r * be.pitch() ) );
    be.ref0 = J.ref + pel; be.ref1 = NULL; be.ref_w = J.ref_w + pel;
    search_block<decltype( be ), FIRST>( be, J, P, tab, bx, by, blk );
}

// Exact raster-serial mode: the reference evaluates blocks in reverse raster order and
// each block reads the MVs of its right / below / below-left / below-right neighbours from the SAME pass
// (slicetype.c:662-680,830-833).  Blocks with equal w = (mbw-1-x) + 2*(mbh-1-y) depend only on blocks of smaller w,
// so one launch per anti-diagonal wave reproduces the serial result exactly; mv_in == mv_out (in place).
__global__ void __launch_bounds__(CTA_THREADS) search_kernel_wave( const SearchJob *__restrict__ jobs, LaParams P, int wave )
{
    const CostTab tab = stage_cost_tab( P );
    SearchJob J = jobs[blockIdx.y];
    resolve_device_weight( J );
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

// Exact raster-serial mode as ONE launch per batch: a warp owns one block row of one field and walks it right to left.
// Block x of a row needs the MVs of the below, below-left and below-right blocks (slicetype.c:662-680).  The field is
// pre-filled with an impossible MV (0x7F7F7F7F); every MV word is stored once, straight to L2 (st.cg), and is its own
// "ready" flag: the row above polls the three words it needs with L1-bypassing loads -- no counters, no fences, and
// the immutable pixel planes keep their L1 residency (an ld.acquire poll invalidated the whole L1: first ncu capture,
// 17 % of the stall samples in CCTL and a 65 % L1 hit rate).  Rows are numbered bottom-up, so a warp only waits for a
// lower-numbered warp, which the hardware dispatched earlier: no deadlock; a bounded spin turns a violated assumption
// into an error flag instead of a hang.
#define X264CU_MV_NOT_READY 0x7F7F7F7F
#ifndef X264CU_SERIAL_PREFETCH
#define X264CU_SERIAL_PREFETCH 2
#endif
#ifndef SERIAL_WARPS
#define SERIAL_WARPS 2
#endif
// candidate slots per round of the exact-mode kernel: 4 (WarpBE, default) or 8 (WarpBE8, tuning build).  Measured on B200
// (profiles/r02/tuning.md): eight slots shorten one stream's serial chain by 4 % (1.213 -> 1.166 ms per step) but cost 4 - 6 % of
// the throughput with 4 - 8 streams per GPU (two rows per lane, eight slots scored even when a round has three candidates).
#ifndef X264CU_SERIAL_SLOTS
#define X264CU_SERIAL_SLOTS 4
#endif
template<bool SATD>
__global__ void __launch_bounds__(SERIAL_WARPS*32) search_kernel_serial( const SearchJob *__restrict__ jobs, LaParams P, int first_job, int njobs,
                                                                          int *__restrict__ err )
{
    const CostTab tab = stage_cost_tab( P );
    const int worker = blockIdx.x * SERIAL_WARPS + ( threadIdx.x >> 5 ), lane = threadIdx.x & 31;
    if( worker >= njobs * P.mbh ) return;
    const int ji = first_job + worker / P.mbh, by = P.mbh - 1 - worker % P.mbh;      // bottom rows first
    SearchJob J = jobs[ji];
    resolve_device_weight( J );
#if X264CU_SERIAL_SLOTS == 8
    WarpBE8<SATD> be; be.q = lane & 3; be.sub = lane >> 2; be.P = &P; be.ref1 = NULL; be.err = err;
    const int row0 = 2 * be.q;
#else
    WarpBE<SATD> be; be.g.r = lane & 7; be.g.mask = 0xffffffffu; be.sub = lane >> 3; be.P = &P; be.ref1 = NULL; be.err = err;
    const int row0 = be.g.r;
#endif
    int right_mv = 0;
    for( int i = 0; i < P.mbw; i++ )
    {
        const int bx = P.mbw - 1 - i, blk = by * P.mbw + bx;
        if( !skipped_border( P, bx, by ) )
        {
            const int64_t pel = (int64_t)( by*8 + row0 ) * P.stride + bx*8;
#if X264CU_SERIAL_SLOTS == 8
            be.fenc0 = *(const uint2 *)( J.fenc + pel ); be.fenc1 = *(const uint2 *)( J.fenc + pel + P.stride );
#else
            be.fenc = *(const uint2 *)( J.fenc + pel );
#endif
            be.ref0 = J.ref + pel; be.ref_w = J.ref_w + pel;
            right_mv = search_block( be, J, P, tab, bx, by, blk, right_mv );
            // warm L1 for the next block of this row: 32 rows x 32 columns of the four phase planes around its
            // position displaced by this block's MV (a miss inside a candidate round costs an L2 round trip on the
            // serial critical path; second ncu capture: 6 % sector misses ~ one per round)
            if( bx > 0 )
            {
                const int cx = ( bx - 1 ) * 8 + ( (int)(int16_t)( right_mv & 0xffff ) >> 2 ) - 12, cy = by * 8 + ( ( right_mv >> 16 ) >> 2 ) - 12 + lane;
                const int y = min( max( cy, -X264CU_PAD ), P.lh + X264CU_PAD - 1 ), x = min( max( cx, -X264CU_PAD ), P.lw + X264CU_PAD - 33 );
                const uint8_t *p = J.ref + (int64_t)y * P.stride + x;
                (void)p;
#pragma unroll
                for( int k = 0; k < 4; k++ )
                {
#if X264CU_SERIAL_PREFETCH == 2
                    // plain loads whose results are never read: the warp does not wait for them (prefetch.global.L1 measured
                    // 16 % slower than no prefetch at all, these 3 % faster; profiles/r01/tuning.md)
                    u32 d0, d1;
                    asm volatile( "ld.global.nc.u8 %0, [%1];" : "=r"( d0 ) : "l"( p + k * P.plane_size ) );
                    asm volatile( "ld.global.nc.u8 %0, [%1];" : "=r"( d1 ) : "l"( p + k * P.plane_size + 31 ) );
                    (void)d0; (void)d1;
#elif X264CU_SERIAL_PREFETCH == 1
                    asm volatile( "prefetch.global.L1 [%0];" :: "l"( p + k * P.plane_size ) );
                    asm volatile( "prefetch.global.L1 [%0];" :: "l"( p + k * P.plane_size + 31 ) );
#endif
                }
            }
        }
        else
        {   // never evaluated (slicetype.c:823-828): the field keeps frame_new's zero MV; publish it for the row above
            right_mv = 0;
            if( lane == 0 ) __stcg( (int *)J.mv_out + blk, 0 );
        }
    }
}

// returns the number of launches
int launch_search_serial( const SearchJob *d_jobs, int njobs, const LaParams &P, int *d_err, int max_resident_ctas, cudaStream_t st )
{
    // every warp of a launch should be resident: split the fields over several launches if one would not fit
    const int warps_per_job = P.mbh, ctas_per_job = ( warps_per_job + SERIAL_WARPS - 1 ) / SERIAL_WARPS;
    int jobs_per_launch = max( 1, max_resident_ctas / max( ctas_per_job, 1 ) ), n = 0;
    for( int j0 = 0; j0 < njobs; n++ )
    {
        // a COOPERATIVE launch: the rows of a launch wait for one another, so every CTA must be resident at the same
        // time.  The runtime refuses a grid that cannot be (and schedules the grid as a whole when other contexts share
        // the device) instead of letting a partly resident launch spin into its time-out.
        int nj = min( jobs_per_launch, njobs - j0 );
        for( ;; )
        {
            const int ctas = ( nj * warps_per_job + SERIAL_WARPS - 1 ) / SERIAL_WARPS;
            int first = j0;
            void *args[] = { (void *)&d_jobs, (void *)&P, (void *)&first, (void *)&nj, (void *)&d_err };
            const void *fn = P.b_satd ? (const void *)search_kernel_serial<true> : (const void *)search_kernel_serial<false>;
            const cudaError_t e = cudaLaunchCooperativeKernel( fn, dim3( ctas ), dim3( SERIAL_WARPS*32 ), args, 0, st );
            if( e == cudaSuccess ) break;
            cudaGetLastError();
            if( e != cudaErrorCooperativeLaunchTooLarge || nj == 1 ) return -1;
            nj = ( nj + 1 ) / 2;
        }
        j0 += nj;
    }
    return n;
}

int search_serial_max_ctas_per_sm()
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor( &n, search_kernel_serial<true>, SERIAL_WARPS*32, 0 );
    return n;
}

void launch_search_waves( const SearchJob *d_jobs, int njobs, const LaParams &P, cudaStream_t st )
{
    dim3 grid( ( P.mbh + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, njobs );
    const int nwaves = P.mbw + 2 * ( P.mbh - 1 );
    for( int w = 0; w < nwaves; w++ )
        search_kernel_wave<<<grid, CTA_THREADS, 0, st>>>( d_jobs, P, w );
}

template<int PITCH>
static void launch_search_t( const SearchJob *d_jobs, int njobs, bool any_weighted, bool first_pass, const LaParams &P, cudaStream_t st )
{
    const int warps_per_cta = S_TILE_W * S_TILE_H / 32;
    dim3 grid( ( P.ntiles_s + warps_per_cta - 1 ) / warps_per_cta, njobs );
    if( P.b_satd )
    {
        if( any_weighted ) search_kernel_t<true, true, PITCH><<<grid, S_TILE_W*S_TILE_H, 0, st>>>( d_jobs, P );
#if X264CU_FIRST_PASS
        else if( first_pass && PITCH ) search_kernel_t<true, false, PITCH, true><<<grid, S_TILE_W*S_TILE_H, 0, st>>>( d_jobs, P );
#endif
        else               search_kernel_t<true, false, PITCH><<<grid, S_TILE_W*S_TILE_H, 0, st>>>( d_jobs, P );
    }
    else
    {
        if( any_weighted ) search_kernel_t<false, true, PITCH><<<grid, S_TILE_W*S_TILE_H, 0, st>>>( d_jobs, P );
        else               search_kernel_t<false, false, PITCH><<<grid, S_TILE_W*S_TILE_H, 0, st>>>( d_jobs, P );
    }
}

// first_pass: every job of the launch has mv_in == NULL (pass 0 of the dependency-free schedule)
void launch_search( const SearchJob *d_jobs, int njobs, bool any_weighted, bool first_pass, const LaParams &P, cudaStream_t st )
{
    if( use_thread_mapping( P, njobs ) )
    {   // the context picks the pitch from {1024, 2048} whenever the padded width fits (capi.cu), anything wider runs the
        // run-time-pitch instantiation
        if( P.stride == X264CU_PITCH_A )      launch_search_t<X264CU_PITCH_A>( d_jobs, njobs, any_weighted, first_pass, P, st );
        else if( P.stride == X264CU_PITCH_B ) launch_search_t<X264CU_PITCH_B>( d_jobs, njobs, any_weighted, first_pass, P, st );
        else                        launch_search_t<0>( d_jobs, njobs, any_weighted, first_pass, P, st );
    }
    else
    {
        dim3 grid( ( P.nmb + GROUPS_PER_CTA - 1 ) / GROUPS_PER_CTA, njobs );
        search_kernel_g<<<grid, CTA_THREADS, 0, st>>>( d_jobs, P );
    }
}
