// slicetype.cpp -- host-side slicetype decision over a frame-cost oracle (plain C++, no CUDA).
//
// Restates the control logic of x264_slicetype_analyse (encoder/slicetype.c:1473-1743): scenecut
// (:1392-1471), the B-adapt trellis slicetype_path / slicetype_path_cost (:1288-1390), B-adapt fast
// (:1580-1622), fixed B placement (:1623-1645), first-minigop scenecut (:1652-1662) and the keyframe
// limit (:1680-1729).  Every slicetype_frame_cost(p0,p1,b) becomes a callback, which the caller backs
// with the device cost matrix (x264cu_frame_cost / x264cu_stream_*) -- or, in the tests, with the CPU
// oracle or with the unmodified reference, to show the three agree frame type by frame type.
// MB-tree and VBV planning (:1677-1678,:1733-1734) consume the same costs and are "next" rows.
#include "../../include/x264cu.h"
#include <cstdint>
#include <cstring>

enum { T_AUTO = 0, T_IDR = 1, T_I = 2, T_P = 3, T_BREF = 4, T_B = 5, T_KEYFRAME = 6 };   // x264.h:278-287
static inline bool is_i( int t ) { return t == T_I || t == T_IDR || t == T_KEYFRAME; }
static inline bool is_b( int t ) { return t == T_B || t == T_BREF; }
static inline bool auto_or_i( int t ) { return t == T_AUTO || is_i( t ); }
static inline bool auto_or_b( int t ) { return t == T_AUTO || is_b( t ); }

#define LOOKAHEAD_MAX 250            // X264_LOOKAHEAD_MAX (common/base.h:140)
#define COST_MAX64 (1ULL<<60)        // encoder/me.h

struct Ctx
{
    const x264cu_slicetype_cfg *cfg;
    x264cu_slicetype_frame *f;       // f[0] = last non-B frame, f[1..n] = lookahead
    x264cu_cost_fn cost; void *user;
    int i_last_keyframe;
    int fc( int p0, int p1, int b ) { int32_t st[4] = { 0, 0, 0, 0 }; return cost( user, p0, p1, b, st ); }
};

// slicetype.c:1288-1331
static uint64_t path_cost( Ctx &c, const char *path, uint64_t threshold, int base )
{
    uint64_t cost = 0;
    int loc = 1, cur_nonb = 0;
    path--;     // the first path element is really the second frame
    while( path[loc] )
    {
        int next_nonb = loc;
        while( path[next_nonb] == 'B' ) next_nonb++;
        if( path[next_nonb] == 'P' ) cost += c.fc( base + cur_nonb, base + next_nonb, base + next_nonb );
        else                         cost += c.fc( base + next_nonb, base + next_nonb, base + next_nonb );
        if( cost > threshold ) break;
        if( c.cfg->bframe_pyramid && next_nonb - cur_nonb > 2 )
        {
            int middle = cur_nonb + ( next_nonb - cur_nonb ) / 2;
            cost += c.fc( base + cur_nonb, base + next_nonb, base + middle );
            for( int nb = loc; nb < middle && cost < threshold; nb++ )
                cost += c.fc( base + cur_nonb, base + middle, base + nb );
            for( int nb = middle+1; nb < next_nonb && cost < threshold; nb++ )
                cost += c.fc( base + middle, base + next_nonb, base + nb );
        }
        else
            for( int nb = loc; nb < next_nonb && cost < threshold; nb++ )
                cost += c.fc( base + cur_nonb, base + next_nonb, base + nb );
        loc = next_nonb + 1;
        cur_nonb = next_nonb;
    }
    return cost;
}

// slicetype.c:1333-1390
static void trellis_path( Ctx &c, int length, char ( *best_paths )[LOOKAHEAD_MAX+1] )
{
    char paths[2][LOOKAHEAD_MAX+1];
    int num_paths = c.cfg->bframes + 1 < length ? c.cfg->bframes + 1 : length;
    uint64_t best_cost = COST_MAX64;
    int best_possible = 0, idx = 0;
    for( int path = 0; path < num_paths; path++ )
    {
        int len = length - ( path + 1 );
        memcpy( paths[idx], best_paths[len % ( X264CU_BFRAME_MAX+1 )], len );
        memset( paths[idx]+len, 'B', path );
        strcpy( paths[idx]+len+path, "P" );
        int possible = 1;
        for( int i = 1; i <= length; i++ )
        {
            int t = c.f[i].type;
            if( t == T_AUTO ) continue;
            if( is_b( t ) )
                possible = possible && ( i < len || i == length || paths[idx][i-1] == 'B' );
            else
            {
                possible = possible && ( i < len || paths[idx][i-1] != 'B' );
                paths[idx][i-1] = is_i( t ) ? 'I' : 'P';
            }
        }
        if( possible || !best_possible )
        {
            if( possible && !best_possible ) best_cost = COST_MAX64;
            uint64_t cost = path_cost( c, paths[idx], best_cost, 0 );
            if( cost < best_cost ) { best_cost = cost; best_possible = possible; idx ^= 1; }
        }
    }
    memcpy( best_paths[length % ( X264CU_BFRAME_MAX+1 )], paths[idx^1], length );
}

// slicetype.c:1392-1428
static int scenecut_internal( Ctx &c, int p0, int p1 )
{
    int32_t st[4] = { 0, 0, 0, 0 }, sti[4] = { 0, 0, 0, 0 };
    int pcost = c.cost( c.user, p0, p1, p1, st );
    int icost = c.cost( c.user, p1, p1, p1, sti );
    const x264cu_slicetype_cfg *g = c.cfg;
    int gop_size = c.f[p1].frame_num - c.i_last_keyframe;
    float thresh_max = g->scenecut_threshold / 100.0;
    float thresh_min = thresh_max * 0.25;
    float bias;
    if( g->keyint_min == g->keyint_max ) thresh_min = thresh_max;
    if( gop_size <= g->keyint_min / 4 || g->intra_refresh ) bias = thresh_min / 4;
    else if( gop_size <= g->keyint_min ) bias = thresh_min * gop_size / g->keyint_min;
    else bias = thresh_min + ( thresh_max - thresh_min ) * ( gop_size - g->keyint_min ) / ( g->keyint_max - g->keyint_min );
    return pcost >= ( 1.0 - bias ) * icost;
}

// slicetype.c:1430-1471
static int scenecut( Ctx &c, int p0, int p1, int real_scenecut, int num_frames, int i_max_search )
{
    const x264cu_slicetype_cfg *g = c.cfg;
    if( real_scenecut && g->bframes )
    {
        int origmaxp1 = p0 + 1;
        if( g->b_adapt == 2 ) origmaxp1 += g->bframes;
        else origmaxp1++;
        int maxp1 = origmaxp1 < num_frames ? origmaxp1 : num_frames;
        for( int curp1 = p1; curp1 <= maxp1; curp1++ )
            if( !scenecut_internal( c, p0, curp1 ) )
                for( int i = curp1; i > p0; i-- ) c.f[i].b_scenecut = 0;
        for( int curp0 = p0; curp0 <= maxp1; curp0++ )
            if( origmaxp1 > i_max_search || ( curp0 < maxp1 && scenecut_internal( c, curp0, maxp1 ) ) )
                c.f[curp0].b_scenecut = 0;
    }
    if( !c.f[p1].b_scenecut ) return 0;
    return scenecut_internal( c, p0, p1 );
}

// x264_slicetype_analyse, slicetype.c:1473-1743 (without MB-tree / VBV lookahead, which only consume costs)
extern "C" int x264cu_slicetype_analyse( const x264cu_slicetype_cfg *cfg, x264cu_slicetype_frame *frames, int n_lookahead,
                                         int i_last_keyframe, int intra_minigop, x264cu_cost_fn cost, void *user )
{
    if( !cfg || !frames || !cost || n_lookahead < 0 || n_lookahead > LOOKAHEAD_MAX ) return -1;
    Ctx c = { cfg, frames, cost, user, i_last_keyframe };
    const int keyframe = !!intra_minigop;
    const int framecnt = n_lookahead, i_max_search = n_lookahead;
    if( !framecnt ) return 0;
    int keyint_limit = cfg->keyint_max - frames[0].frame_num + i_last_keyframe - 1;
    int orig_num_frames, num_frames;
    orig_num_frames = num_frames = cfg->intra_refresh ? framecnt : ( framecnt < keyint_limit ? framecnt : keyint_limit );
    if( cfg->psy_mbtree ) num_frames = framecnt;                         // :1511-1512
    else if( cfg->open_gop && num_frames < framecnt ) num_frames++;
    else if( num_frames == 0 ) { frames[1].type = T_I; return 0; }

    if( auto_or_i( frames[1].type ) && cfg->scenecut_threshold && scenecut( c, 0, 1, 1, orig_num_frames, i_max_search ) )
    {
        if( frames[1].type == T_AUTO ) frames[1].type = T_I;
        return 0;
    }
    for( int j = 1; j <= num_frames; j++ )
        if( frames[j].type == T_KEYFRAME ) frames[j].type = cfg->open_gop ? T_I : T_IDR;
    for( int j = 2; j <= num_frames; j++ )
        if( frames[j].type == T_IDR && auto_or_b( frames[j-1].type ) ) frames[j-1].type = T_P;

    int num_analysed_frames = num_frames, reset_start;
    if( cfg->bframes )
    {
        if( cfg->b_adapt == 2 )
        {
            if( num_frames > 1 )
            {
                static thread_local char best_paths[X264CU_BFRAME_MAX+1][LOOKAHEAD_MAX+1];
                memset( best_paths, 0, sizeof(best_paths) );
                strcpy( best_paths[1], "P" );
                int best_path_index = num_frames % ( X264CU_BFRAME_MAX+1 );
                for( int j = 2; j <= num_frames; j++ ) trellis_path( c, j, best_paths );
                for( int j = 1; j < num_frames; j++ )
                {
                    if( best_paths[best_path_index][j-1] != 'B' ) { if( auto_or_b( frames[j].type ) ) frames[j].type = T_P; }
                    else if( frames[j].type == T_AUTO ) frames[j].type = T_B;
                }
            }
        }
        else if( cfg->b_adapt == 1 )
        {
            int last_nonb = 0, num_bframes = cfg->bframes;
            char path[LOOKAHEAD_MAX+1];
            for( int j = 1; j < num_frames; j++ )
            {
                if( j-1 > 0 && is_b( frames[j-1].type ) ) num_bframes--;
                else { last_nonb = j-1; num_bframes = cfg->bframes; }
                if( !num_bframes ) { if( auto_or_b( frames[j].type ) ) frames[j].type = T_P; continue; }
                if( frames[j].type != T_AUTO ) continue;
                if( is_b( frames[j+1].type ) ) { frames[j].type = T_P; continue; }
                int bframes = j - last_nonb - 1;
                memset( path, 'B', bframes );
                strcpy( path+bframes, "PP" );
                uint64_t cost_p = path_cost( c, path, COST_MAX64, last_nonb );
                strcpy( path+bframes, "BP" );
                uint64_t cost_b = path_cost( c, path, cost_p, last_nonb );
                frames[j].type = cost_b < cost_p ? T_B : T_P;
            }
        }
        else
        {
            int num_bframes = cfg->bframes;
            for( int j = 1; j < num_frames; j++ )
            {
                if( !num_bframes ) { if( auto_or_b( frames[j].type ) ) frames[j].type = T_P; }
                else if( frames[j].type == T_AUTO ) frames[j].type = is_b( frames[j+1].type ) ? T_P : T_B;
                if( is_b( frames[j].type ) ) num_bframes--;
                else num_bframes = cfg->bframes;
            }
        }
        if( auto_or_b( frames[num_frames].type ) ) frames[num_frames].type = T_P;
        int num_bframes = 0;
        while( num_bframes < num_frames && is_b( frames[num_bframes+1].type ) ) num_bframes++;
        for( int j = 1; j < num_bframes+1; j++ )
            if( frames[j].forced_type == T_AUTO && auto_or_i( frames[j+1].forced_type ) && cfg->scenecut_threshold &&
                scenecut( c, j, j+1, 0, orig_num_frames, i_max_search ) )
            {
                frames[j].type = T_P;
                num_analysed_frames = j;
                break;
            }
        reset_start = keyframe ? 1 : ( num_bframes+2 < num_analysed_frames+1 ? num_bframes+2 : num_analysed_frames+1 );
    }
    else
    {
        for( int j = 1; j <= num_frames; j++ )
            if( auto_or_b( frames[j].type ) ) frames[j].type = T_P;
        reset_start = !keyframe + 1;
    }

    if( !cfg->intra_refresh )
    {   // enforce the keyframe limit, slicetype.c:1680-1729
        int last_keyframe = i_last_keyframe, last_possible = 0;
        for( int j = 1; j <= num_frames; j++ )
        {
            x264cu_slicetype_frame *frm = &frames[j];
            int keyframe_dist = frm->frame_num - last_keyframe;
            if( auto_or_i( frm->forced_type ) )
                if( cfg->open_gop || !is_b( frames[j-1].forced_type ) ) last_possible = j;
            if( keyframe_dist >= cfg->keyint_max )
            {
                if( last_possible != 0 && last_possible != j )
                {
                    j = last_possible;
                    frm = &frames[j];
                    keyframe_dist = frm->frame_num - last_keyframe;
                }
                last_possible = 0;
                if( frm->type != T_IDR ) frm->type = cfg->open_gop ? T_I : T_IDR;
            }
            if( frm->type == T_I && keyframe_dist >= cfg->keyint_min )
            {
                if( cfg->open_gop ) last_keyframe = frm->frame_num;
                else if( frm->forced_type != T_I ) frm->type = T_IDR;
            }
            if( frm->type == T_IDR )
            {
                last_keyframe = frm->frame_num;
                if( j > 1 && is_b( frames[j-1].type ) ) frames[j-1].type = T_P;
            }
        }
    }
    // frames that have not actually been decided yet get their forced type back (:1736-1738); the caller reads
    // decided[] = the types before this reset for j < reset_start, and the analysis result for the rest in `analysed`
    for( int j = 1; j <= num_frames; j++ ) frames[j].analysed_type = frames[j].type;
    for( int j = reset_start; j <= num_frames; j++ ) frames[j].type = frames[j].forced_type;
    return reset_start;
}
