// mbtree_host.cpp -- the frame walk of macroblock_tree (encoder/slicetype.c:1091-1188) over the C ABI.
// Host control only: every slicetype_frame_cost becomes x264cu_frame_cost (cache hits after the lookahead
// has evaluated the matrix) and every macroblock_tree_propagate / _finish a device launch.
#include "../../include/x264cu.h"
#include <cmath>
#include <cstring>

static inline bool is_b( int t ) { return t == 4 || t == 5; }          // X264_TYPE_BREF / X264_TYPE_B
static inline float clip_duration( float f ) { return f < 0.01f ? 0.01f : f > 1.00f ? 1.00f : f; }   // ratecontrol.h:37-40
#define MBTREE_PRECISION 0.5f                                           // slicetype.c:1027

struct Walk
{
    x264cu_ctx *ctx; const x264cu_mbtree_cfg *cfg; x264cu_mbtree_frame *fr; float average_duration; int err;
    void cost( int p0, int p1, int b )
    {   // slicetype_frame_cost( h, a, frames, p0, p1, b ): only its side effects on the device caches are needed here
        x264cu_job j;
        memset( &j, 0, sizeof(j) );
        j.fenc = fr[b].f; j.fref0 = fr[p0].f; j.fref1 = fr[p1].f; j.d0 = b - p0; j.d1 = p1 - b;
        if( x264cu_frame_cost( ctx, &j ) ) err = -1;
    }
    void reset( int i ) { if( x264cu_mbtree_reset( ctx, fr[i].f ) ) err = -1; }
    void finish( int i, int ref0_distance )
    {   // macroblock_tree_finish, slicetype.c:1029-1049
        int fps_factor = (int)round( clip_duration( average_duration ) / clip_duration( fr[i].duration ) * 256 / MBTREE_PRECISION );
        float weightdelta = 0.0;
        if( ref0_distance && fr[i].weighted_cost_delta[ref0_distance-1] > 0 )
            weightdelta = ( 1.0 - fr[i].weighted_cost_delta[ref0_distance-1] );
        float strength = 5.0f * ( 1.0f - cfg->qcompress );
        if( x264cu_mbtree_finish( ctx, fr[i].f, fps_factor, weightdelta, strength, fr[i].qp_offset_aq, fr[i].qp_offset ) ) err = -1;
    }
    void propagate( int p0, int p1, int b, int referenced )
    {   // macroblock_tree_propagate, slicetype.c:1051-1089
        float fps_factor = clip_duration( fr[b].duration ) / ( clip_duration( average_duration ) * 256.0f ) * MBTREE_PRECISION;
        if( x264cu_mbtree_propagate( ctx, fr[b].f, fr[p0].f, fr[p1].f, b - p0, p1 - b, fps_factor, referenced ) ) err = -1;
        if( cfg->vbv_lookahead && referenced ) finish( b, b == p1 ? b - p0 : 0 );
    }
};

extern "C" int x264cu_macroblock_tree( x264cu_ctx *ctx, const x264cu_mbtree_cfg *cfg, x264cu_mbtree_frame *frames, int num_frames, int b_intra )
{
    if( !ctx || !cfg || !frames || num_frames < 0 ) return -1;
    Walk w = { ctx, cfg, frames, 0.f, 0 };
    int mb_count;
    { int info[6]; if( x264cu_info( ctx, info ) ) return -1; mb_count = info[0] * info[1]; }
    int idx = !b_intra, last_nonb, cur_nonb = 1, bframes = 0;
    float total_duration = 0.0;
    for( int j = 0; j <= num_frames; j++ ) total_duration += frames[j].duration;
    w.average_duration = total_duration / ( num_frames + 1 );
    int i = num_frames;
    if( b_intra ) w.cost( 0, 0, 0 );
    while( i > 0 && is_b( frames[i].type ) ) i--;
    last_nonb = i;
    if( !cfg->rc_lookahead )
    {   // lookaheadless MB-tree, slicetype.c:1120-1130
        if( b_intra )
        {
            w.reset( 0 );
            memcpy( frames[0].qp_offset, frames[0].qp_offset_aq, mb_count * sizeof(float) );
            return w.err;
        }
        if( x264cu_mbtree_swap( ctx, frames[last_nonb].f, frames[0].f ) ) return -1;
        w.reset( 0 );
    }
    else
    {
        if( last_nonb < idx ) return w.err;
        w.reset( last_nonb );
    }
    while( i-- > idx )
    {
        cur_nonb = i;
        while( is_b( frames[cur_nonb].type ) && cur_nonb > 0 ) cur_nonb--;
        if( cur_nonb < idx ) break;
        w.cost( cur_nonb, last_nonb, last_nonb );
        w.reset( cur_nonb );
        bframes = last_nonb - cur_nonb - 1;
        if( cfg->bframe_pyramid && bframes > 1 )
        {
            int middle = ( bframes + 1 ) / 2 + cur_nonb;
            w.cost( cur_nonb, last_nonb, middle );
            w.reset( middle );
            while( i > cur_nonb )
            {
                int p0 = i > middle ? middle : cur_nonb;
                int p1 = i < middle ? middle : last_nonb;
                if( i != middle )
                {
                    w.cost( p0, p1, i );
                    w.propagate( p0, p1, i, 0 );
                }
                i--;
            }
            w.propagate( cur_nonb, last_nonb, middle, 1 );
        }
        else
        {
            while( i > cur_nonb )
            {
                w.cost( cur_nonb, last_nonb, i );
                w.propagate( cur_nonb, last_nonb, i, 0 );
                i--;
            }
        }
        w.propagate( cur_nonb, last_nonb, last_nonb, 1 );
        last_nonb = cur_nonb;
    }
    if( !cfg->rc_lookahead )
    {
        w.cost( 0, last_nonb, last_nonb );
        w.propagate( 0, last_nonb, last_nonb, 1 );
        if( x264cu_mbtree_swap( ctx, frames[last_nonb].f, frames[0].f ) ) return -1;
    }
    w.finish( last_nonb, last_nonb );
    if( cfg->bframe_pyramid && bframes > 1 && !cfg->vbv )
        w.finish( last_nonb + ( bframes + 1 ) / 2, 0 );
    return w.err;
}
