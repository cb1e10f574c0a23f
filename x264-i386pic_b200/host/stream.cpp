// stream.cpp -- host-side lookahead window over the C ABI (plain C++, no CUDA here).
//
// One x264cu_stream is what one encoder's lookahead thread holds (encoder/lookahead.c:68-129): a
// sliding window of frames.  Pushing a frame does what x264_frame_init_lowres +
// x264_opencl_slicetype_prep (slicetype-cl.c:653-712) do for the newest frame n, generalised to
// the whole B-frame cost matrix: every slicetype_frame_cost(p0, p1 = n, b) with
// n - (bframes+1) <= p0 < b <= n that slicetype_path (slicetype.c:1333-1390) or scenecut (:1430)
// can ask for, plus the I cost -- 33 new searches and 153 + 1 evaluations per frame at bframes = 16.
// The decision logic then reads costs from the matrix instead of calling into the device.
#include "../../include/x264cu.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct x264cu_stream
{
    x264cu_ctx *ctx;
    int bframes, depth;                 // depth = bframes + 3 frame handles: the window (bframes + 2) + one frame uploaded ahead
    int mbw, mbh, nmb;
    int want_arrays;                    // also bring lowres_costs / row_satds / MVs back (the full hook protocol)
    std::vector<x264cu_frame *> ring;   // frame n lives in ring[n % depth]
    int64_t n_pushed;
    uint64_t last_seq;                  // batch sequence number of the most recent push (x264cu_wait target)
    // host results of the most recent pushes: stats[d0][d1][4] for p1 = newest.  One set per push, NSETS deep: the
    // deferred copies of push n land when its batch is awaited, which may be after push n+1 was enqueued, so a push
    // never writes into (or resets) the set of the one before it.  A context keeps at most 4 batches = 2 pushes in flight.
    enum { NSETS = 4 };
    std::vector<int32_t> stats_set[NSETS];      // (B+2)*(B+2)*4, indexed [d0*(B+2)+d1]
    std::vector<int32_t> intra_stats;   // NSETS x 2
    int32_t *stats_of( int64_t push ) { return stats_set[push % NSETS].data(); }
    // optional array destinations (pinned when possible)
    uint8_t *arr; size_t arr_bytes; bool arr_pinned;
    std::vector<x264cu_job> jobs;
    std::vector<x264cu_out> eager;      // want_arrays: where the arrays of evaluation [d0*(B+2)+d1] of the most recent push landed
    int shard_ahead;                    // sharded pushes whose phase 0 ran but whose phase 1 did not yet (0..2)
};

extern "C" void *x264cu_host_alloc( x264cu_ctx *ctx, size_t bytes );
extern "C" void  x264cu_host_free( x264cu_ctx *ctx, void *p );

extern "C" x264cu_stream *x264cu_stream_open( x264cu_ctx *ctx, int bframes, int want_arrays )
{
    int info[6];
    if( !ctx || bframes < 0 || bframes > X264CU_BFRAME_MAX || x264cu_info( ctx, info ) ) return NULL;
    x264cu_stream *s = new x264cu_stream();
    s->ctx = ctx; s->bframes = bframes; s->depth = bframes + 3;
    s->mbw = info[0]; s->mbh = info[1]; s->nmb = info[0] * info[1];
    s->want_arrays = want_arrays;
    s->n_pushed = 0; s->last_seq = 0; s->shard_ahead = 0;
    for( auto &v : s->stats_set ) v.assign( (size_t)( bframes+2 ) * ( bframes+2 ) * 4, -1 );
    s->intra_stats.assign( 2 * x264cu_stream::NSETS, -1 );
    s->arr = NULL; s->arr_bytes = 0; s->arr_pinned = false;
    for( int i = 0; i < s->depth; i++ )
    {
        x264cu_frame *f = x264cu_frame_new( ctx );
        if( !f ) { for( x264cu_frame *g : s->ring ) x264cu_frame_free( ctx, g ); delete s; return NULL; }
        s->ring.push_back( f );
    }
    if( want_arrays )
    {   // per evaluation: lowres_costs (u16 x nmb) + row_satds (i32 x mbh); per new search: mvs (i16x2 x nmb)
        size_t ntri = (size_t)( bframes+1 ) * ( bframes+2 ) / 2 + 1, nsearch = 2 * (size_t)bframes + 1;
        s->arr_bytes = ntri * ( (size_t)s->nmb * 2 + (size_t)s->mbh * 4 + 64 ) + nsearch * ( (size_t)s->nmb * 4 + 64 );
        s->arr = (uint8_t *)x264cu_host_alloc( ctx, s->arr_bytes );
        s->arr_pinned = s->arr != NULL;
        if( !s->arr ) s->arr = (uint8_t *)malloc( s->arr_bytes );
    }
    return s;
}

extern "C" void x264cu_stream_close( x264cu_stream *s )
{
    if( !s ) return;
    x264cu_flush( s->ctx );
    for( x264cu_frame *f : s->ring ) x264cu_frame_free( s->ctx, f );
    if( s->arr ) { if( s->arr_pinned ) x264cu_host_free( s->ctx, s->arr ); else free( s->arr ); }
    delete s;
}

// Enqueue frame n = number of frames pushed so far.  on_device: `luma`/`inv_qscale` are device pointers.
// Returns the number of evaluations enqueued, or -1.
extern "C" int x264cu_stream_push( x264cu_stream *s, const uint8_t *luma, int stride, const uint16_t *inv_qscale, int on_device )
{
    if( !s || !luma ) return -1;
    const int B = s->bframes, D = s->depth;
    const int64_t n = s->n_pushed;
    x264cu_frame *fn = s->ring[n % D];
    // the eager result arrays of the previous push are copied straight into s->arr: finish it before re-using them
    if( s->want_arrays && s->last_seq && x264cu_wait( s->ctx, s->last_seq ) ) return -1;
    int r = on_device ? x264cu_frame_upload_device( s->ctx, fn, luma, stride, inv_qscale )
                      : x264cu_frame_upload( s->ctx, fn, luma, stride, inv_qscale );
    if( r ) return -1;
    s->jobs.clear();
    int32_t *stats = s->stats_of( n );
    std::fill( stats, stats + s->stats_set[0].size(), -1 );
    s->eager.assign( (size_t)( B+2 ) * ( B+2 ), x264cu_out() );
    uint8_t *ap = s->arr;
    auto take = [&]( size_t bytes ) { uint8_t *p = ap; ap += ( bytes + 63 ) & ~(size_t)63; return p; };
    // the I cost of the new frame (scenecut needs icost, slicetype.c:1401)
    {
        x264cu_job j;
        memset( &j, 0, sizeof(j) );
        j.fenc = j.fref0 = j.fref1 = fn; j.d0 = j.d1 = 0;
        j.out.stats = &stats[0];
        j.out.intra_stats = &s->intra_stats[2 * ( n % x264cu_stream::NSETS )];
        if( s->want_arrays ) j.out.intra_costs = (uint16_t *)take( (size_t)s->nmb * 2 );
        s->eager[0] = j.out;
        s->jobs.push_back( j );
    }
    // P evaluations first (they provide fref1->lowres_mvs[0] for the temporal-direct candidate of the
    // B evaluations between p0 and n, slicetype.c:629), then the B evaluations, as slicetype_path_cost orders them
    const int64_t lo = n - ( B + 1 ) < 0 ? 0 : n - ( B + 1 );
    for( int64_t p0 = n - 1; p0 >= lo; p0-- )
        for( int k = 0; k < (int)( n - p0 ); k++ )
        {   // k == 0: the P evaluation b = n; then the B evaluations b = p0+1 .. n-1
            const int64_t b = k == 0 ? n : p0 + k;
            x264cu_job j;
            memset( &j, 0, sizeof(j) );
            j.fenc = s->ring[b % D]; j.fref0 = s->ring[p0 % D]; j.fref1 = fn;
            j.d0 = (int)( b - p0 ); j.d1 = (int)( n - b );
            j.out.stats = &stats[( (size_t)j.d0 * ( B+2 ) + j.d1 ) * 4];
            if( s->want_arrays )
            {
                j.out.lowres_costs = (uint16_t *)take( (size_t)s->nmb * 2 );
                j.out.row_satds = (int32_t *)take( (size_t)s->mbh * 4 );
                // MV fields of the searches that are new with frame n: L0 of frame n, L1 of older frames towards n
                if( b == n ) j.out.mvs[0] = (int16_t *)take( (size_t)s->nmb * 4 );
                else if( p0 == b - 1 ) j.out.mvs[1] = (int16_t *)take( (size_t)s->nmb * 4 );
            }
            s->eager[(size_t)j.d0 * ( B+2 ) + j.d1] = j.out;
            s->jobs.push_back( j );
        }
    if( x264cu_frame_cost_batch( s->ctx, s->jobs.data(), (int)s->jobs.size() ) ) return -1;
    s->last_seq = x264cu_last_batch( s->ctx );
    s->n_pushed++;
    return (int)s->jobs.size();
}

// ---- one stream over several GPUs (BASELINE.json configs[3]) -------------------------------------
// Every rank holds the whole window (it uploads every frame; lowres + intra are replicated, they are 5 % of a step) and
// owns a share of the step's units: the list-0 search (n -> p0) and the P evaluation (p0, n, n) belong to rank p0 % world,
// the list-1 search (b -> n) and the B evaluations (p0, n, b) to rank b % world, the I evaluation to rank n % world.
// The only data another rank needs are the list-0 fields of the newest frame (temporal-direct source of this step's B
// evaluations, list-0 side of later ones): the caller all-gathers them between phase 0 and phase 1.
extern "C" x264cu_frame *x264cu_stream_frame( x264cu_stream *s, int64_t index )
{   // frame number `index` (0 = the first frame pushed); valid for the frames of the current window and the one uploaded ahead
    if( !s || index < 0 || index > s->n_pushed + 1 || index < s->n_pushed - ( s->bframes + 1 ) ) return NULL;
    return s->ring[index % s->depth];
}

extern "C" int x264cu_stream_push_shard( x264cu_stream *s, const uint8_t *luma, int stride, const uint16_t *inv_qscale, int on_device,
                                         int rank, int world, int phase )
{
    if( !s || world < 1 || rank < 0 || rank >= world ) return -1;
    const int B = s->bframes, D = s->depth;
    // phase 0 may run one frame ahead of phase 1 (the caller overlaps the exchange of step n with the searches of step n+1)
    const int64_t n = phase == 0 ? s->n_pushed + s->shard_ahead : s->n_pushed;
    x264cu_frame *fn = s->ring[n % D];
    const int64_t lo = n - ( B + 1 ) < 0 ? 0 : n - ( B + 1 );
    if( phase == 0 )
    {
        if( !luma || s->shard_ahead >= 2 ) return -1;
        int r = on_device ? x264cu_frame_upload_device( s->ctx, fn, luma, stride, inv_qscale )
                          : x264cu_frame_upload( s->ctx, fn, luma, stride, inv_qscale );
        if( r ) return -1;
        std::vector<x264cu_search_job> sj;
        for( int64_t p0 = n - 1; p0 >= lo; p0-- )
            if( p0 % world == rank )
            {
                x264cu_search_job j; memset( &j, 0, sizeof(j) );
                j.fenc = fn; j.fref = s->ring[p0 % D]; j.list = 0; j.dist = (int)( n - p0 );
                sj.push_back( j );
            }
        for( int64_t b = lo + 1; b < n; b++ )
            if( b % world == rank )
            {
                x264cu_search_job j; memset( &j, 0, sizeof(j) );
                j.fenc = s->ring[b % D]; j.fref = fn; j.list = 1; j.dist = (int)( n - b );
                sj.push_back( j );
            }
        if( x264cu_search_batch( s->ctx, sj.data(), (int)sj.size() ) ) return -1;
        s->shard_ahead++;
        return (int)sj.size();
    }
    if( s->shard_ahead < 1 ) return -1;
    s->jobs.clear();
    int32_t *stats = s->stats_of( n );
    std::fill( stats, stats + s->stats_set[0].size(), -1 );
    if( n % world == rank )
    {
        x264cu_job j; memset( &j, 0, sizeof(j) );
        j.fenc = j.fref0 = j.fref1 = fn;
        j.out.stats = &stats[0];
        j.out.intra_stats = &s->intra_stats[2 * ( n % x264cu_stream::NSETS )];
        s->jobs.push_back( j );
    }
    for( int64_t p0 = n - 1; p0 >= lo; p0-- )
        for( int k = 0; k < (int)( n - p0 ); k++ )
        {
            const int64_t b = k == 0 ? n : p0 + k;
            if( ( k == 0 ? p0 : b ) % world != rank ) continue;
            x264cu_job j; memset( &j, 0, sizeof(j) );
            j.fenc = s->ring[b % D]; j.fref0 = s->ring[p0 % D]; j.fref1 = fn;
            j.d0 = (int)( b - p0 ); j.d1 = (int)( n - b );
            j.out.stats = &stats[( (size_t)j.d0 * ( B+2 ) + j.d1 ) * 4];
            s->jobs.push_back( j );
        }
    if( x264cu_frame_cost_batch( s->ctx, s->jobs.data(), (int)s->jobs.size() ) ) return -1;
    s->last_seq = x264cu_last_batch( s->ctx );
    s->n_pushed++;
    s->shard_ahead--;
    return (int)s->jobs.size();
}

// completes this window's most recent push only; pushes of other windows enqueued after it keep running
extern "C" int x264cu_stream_sync( x264cu_stream *s ) { return s ? x264cu_wait( s->ctx, s->last_seq ) : -1; }

// stats of slicetype_frame_cost( p0 = n-d0-d1.., ) for the newest frame n as p1: d0 = b-p0, d1 = n-b.
// out4: cost_est (B: scaled), cost_est_aq, intra_mbs, raw cost; -1 when not evaluated.
extern "C" int x264cu_stream_cost( x264cu_stream *s, int d0, int d1, int32_t *out4 )
{
    if( !s || d0 < 0 || d1 < 0 || d0 + d1 > s->bframes + 1 ) return -1;
    if( s->n_pushed < 1 ) { out4[0] = out4[1] = out4[2] = out4[3] = -1; return 0; }
    memcpy( out4, &s->stats_of( s->n_pushed - 1 )[( (size_t)d0 * ( s->bframes+2 ) + d1 ) * 4], 16 );
    return 0;
}

// want_arrays: host pointers of the eagerly copied arrays of evaluation (d0, d1) of the most recent push (NULL members:
// not produced -- each new MV field is delivered once, with the first evaluation that uses it); valid after sync and
// until the next push of this window
extern "C" int x264cu_stream_arrays( x264cu_stream *s, int d0, int d1, x264cu_out *out )
{
    if( !s || !out || !s->want_arrays || d0 < 0 || d1 < 0 || d0 + d1 > s->bframes + 1 || s->eager.empty() ) return -1;
    *out = s->eager[(size_t)d0 * ( s->bframes+2 ) + d1];
    return 0;
}

// On-demand result arrays.  The per-evaluation arrays of every frame of the window stay in device memory (the frame's
// cost-matrix storage); a consumer asks for the few it reads -- the decided path of a mini-GOP needs, per frame,
// lowres_costs[b-p0][p1-b], the row SATDs and the MV fields towards its two references (slicetype.c:1056,1078-1083;
// mvpred.c:553) -- instead of all 154 evaluations of every step crossing PCIe.  `b` is a frame number (0 = the first
// frame pushed); b-d0 .. b+d1 must still be in the window.  The evaluation is computed first if it was never asked for
// (slicetype_frame_cost semantics), otherwise this is a cache hit: one gather launch + one D2H transfer.  The arrays
// are complete after x264cu_stream_sync (or x264cu_flush).
extern "C" int x264cu_stream_fetch( x264cu_stream *s, int64_t b, int d0, int d1, const x264cu_out *out )
{
    if( !s || !out || d0 < 0 || d1 < 0 || d0 + d1 > s->bframes + 1 ) return -1;
    const int64_t newest = s->n_pushed - 1, p0 = b - d0, p1 = b + d1;
    if( newest < 0 || p1 > newest || p0 < 0 || p0 < newest - ( s->bframes + 1 ) ) return -1;
    x264cu_job j;
    memset( &j, 0, sizeof(j) );
    j.fenc = s->ring[b % s->depth]; j.fref0 = s->ring[p0 % s->depth]; j.fref1 = s->ring[p1 % s->depth];
    j.d0 = d0; j.d1 = d1;
    j.out = *out;
    if( x264cu_frame_cost_batch( s->ctx, &j, 1 ) ) return -1;
    s->last_seq = x264cu_last_batch( s->ctx );
    return 0;
}

extern "C" int64_t x264cu_stream_frames( x264cu_stream *s ) { return s ? s->n_pushed : -1; }
extern "C" int64_t x264cu_stream_result_bytes( x264cu_stream *s ) { return s ? (int64_t)s->arr_bytes : -1; }
