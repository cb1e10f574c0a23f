// capi.cu -- host side of libx264cuda.so: the C ABI of include/x264cu.h.
//
// Owns what common/opencl.c + encoder/slicetype-cl.c own in the reference: the device context and
// its single in-order stream (opencl.c:525), the page-locked staging buffer and the deferred
// copy list drained by flush (opencl.h:712-718, slicetype-cl.c:58-70), per-frame device objects
// created on first use (slicetype-cl.c:101-160), and the evaluation caches that mirror
// i_cost_est / the 0x7FFF MV sentinel (slicetype.c:848-867).  No CPU fallback anywhere.
#include "../../include/x264cu.h"
#include "lookahead.cuh"
#include "primitives.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

static thread_local char g_err[512] = "";
#define FAIL( ... ) do { snprintf( g_err, sizeof(g_err), __VA_ARGS__ ); return -1; } while( 0 )
#define CU( call ) do { cudaError_t e_ = (call); if( e_ != cudaSuccess ) { \
    snprintf( g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString( e_ ) ); return -1; } } while( 0 )
#define CUP( call ) do { cudaError_t e_ = (call); if( e_ != cudaSuccess ) { \
    snprintf( g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString( e_ ) ); return NULL; } } while( 0 )

enum { K_LOWRES, K_INTRA, K_SEARCH, K_FINAL, K_NCLASS };

struct PendingCopy { void *dst; const void *src; size_t bytes; int kind, bias; uint64_t seq; };   // kind < 0: raw copy; 0/1/2: I/P/B stats fix-up
#define NSLOT 4      // batches that may be in flight before the host has to wait (pipeline depth)
struct TimedSpan { cudaEvent_t a, b; int cls; };

struct x264cu_frame
{
    uint8_t  *planes;        // padded origin of plane F; H,V,C follow at plane_size
    uint16_t *inv_qscale;
    uint16_t *costs;         // [(B+2)*(B+2)][nmb]; [0] is i_intra_cost (alias of lowres_costs[0][0], frame.c:287)
    int16_t  *mvs;           // [2*(B+1)][nmb][2]
    int32_t  *mvcosts;       // [2*(B+1)][nmb]
    int32_t  *stats;         // [(B+2)*(B+2)][stat_pitch]: row_satds[mbh], cost_est, cost_est_aq, intra_mbs
    uint32_t *prop;          // i_propagate_cost accumulators (MB-tree), clamped to 32767 on read
    float    *qp_aq, *qp;    // f_qp_offset_aq (in) / f_qp_offset (out) staging for macroblock_tree_finish
    unsigned long long *pixstats;   // raw {sum, ssd} of Y, U, V (x264cu_frame_upload_aq); [6..7] unused
    int32_t  *wres;          // [bframes+1][8]: device-side result of x264cu_weights_analyse per list-0 distance
    bool      have_aq;       // pixstats / qp_aq were produced by the device AQ kernel
    uint8_t   wanalysed[X264CU_BFRAME_MAX+1];   // wres[dist-1] holds a result
    uint8_t   searched[2][X264CU_BFRAME_MAX+1];
    uint8_t   evaluated[X264CU_BFRAME_MAX+2][X264CU_BFRAME_MAX+2];
    bool      uploaded;
    bool      any_eval;      // some evaluation already produced i_cost_est[0][0] (slicetype.c:848,949-953)
};

struct x264cu_ctx
{
    int device;
    x264cu_cfg cfg;
    LaParams P;
    int sm_count;
    cudaStream_t st;                 // compute stream (the in-order queue every kernel runs on)
    cudaStream_t st_in, st_out;      // copy streams: luma H2D ahead of the compute, results D2H behind it
    cudaEvent_t ev_up[2], ev_res;
    uint16_t *d_cost_mv;
    float *d_log2_lut;
    uint8_t *d_exp2_lut;
    uint32_t *d_pow_x; float *d_pow_v; int pow_n;      // aq-mode 2 / 3: inputs on which libm's powf(x, 0.125f) is not the rounded triple square root
    uint32_t *d_tiles;
    bool exact_waves; int serial_ctas_per_sm; int serial_error;      // exact raster-serial mode (search_passes == 0)
    uint8_t *d_chroma[2]; int chroma_pitch;      // NV12 chroma staging, paired with d_luma[]
    int stat_pitch;
    size_t plane_alloc;      // bytes of the 4 planes + slack
    // full-resolution luma staging on the device (double buffered)
    uint8_t *d_luma[2]; int luma_pitch; int luma_sel;
    cudaEvent_t luma_free[2];
    // pinned staging (H2D inputs, D2H results) and the deferred copy list
    // Every enqueueing entry point is one "batch"; batch k uses region k % NSLOT of each pool below, and a
    // region is recycled only after the CUDA event of the batch that used it last has completed and its
    // deferred copies have been drained -- so up to NSLOT batches are in flight while the host keeps enqueueing.
    uint8_t *pin; size_t pin_size, pin_used[NSLOT];
    std::vector<PendingCopy> copies; size_t copies_head;
    // job tables: pinned host ring mirrored to a device ring
    uint8_t *h_jobs, *d_jobs; size_t jobs_size, jobs_used[NSLOT];
    // device workspace arena for per-batch temporaries (MV ping-pong fields, weighted planes, result gather)
    uint8_t *arena; size_t arena_size, arena_used[NSLOT];
    uint64_t seq; int slot; bool batch_out;      // batch_out: this batch ends on the D2H stream
    cudaEvent_t slot_done[NSLOT]; uint64_t slot_seq[NSLOT]; bool slot_pending[NSLOT];
    int64_t launches;
    std::vector<uint8_t *> batch_flags;      // cache flags set by the batch being built (rolled back if it fails)
    // timing
    bool profiling;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
    float class_ms[K_NCLASS]; int class_n[K_NCLASS];
    cudaEvent_t tm_a, tm_b;
    std::vector<x264cu_frame *> frames;
};

extern "C" const char *x264cu_last_error( void ) { return g_err; }

extern "C" int x264cu_device_count( void )
{
    int n = 0;
    if( cudaGetDeviceCount( &n ) != cudaSuccess ) { cudaGetLastError(); return 0; }
    return n;
}

static int ctx_sync( x264cu_ctx *c );

static inline size_t align_up( size_t v, size_t a ) { return ( v + a - 1 ) / a * a; }

// Warp tiles of the thread-per-block kernels.  The frame is cut into strips of `strip_h` block rows; inside a strip,
// columns are covered left to right by chunks of `tile_w` columns, each chunk by warp tiles of tile_w x (32/tile_w)
// blocks stacked vertically (consecutive table entries = one CTA tile of tile_w x strip_h blocks, as before).  The
// columns left over at the right edge are decomposed into narrower power-of-two widths whose tiles are taller, so that
// every warp still owns 32 real blocks: at 1080p (120 columns) the 32-wide search tiles leave 24 = 16 + 8 columns, which
// become 16x2 and 8x4 tiles instead of a quarter-empty 32x1 tile per row (6 % of all lanes).
static std::vector<uint32_t> build_warp_tiles( int mbw, int mbh, int tile_w, int strip_h )
{
    std::vector<uint32_t> t;
    auto lg = []( int v ) { int l = 0; while( ( 1 << l ) < v ) l++; return l; };
    for( int y0 = 0; y0 < mbh; y0 += strip_h )
    {
        const int rows = std::min( strip_h, mbh - y0 );
        int x = 0;
        while( x < mbw )
        {
            int w = tile_w;
            while( w > 1 && x + w > mbw && 32 / ( w / 2 ) <= strip_h ) w /= 2;      // shrink at the right edge while the taller tile still fits the strip
            const int h = 32 / w;
            for( int y = 0; y < rows; y += h ) t.push_back( (uint32_t)x | ( (uint32_t)( y0 + y ) << 12 ) | ( (uint32_t)lg( w ) << 24 ) );
            x += w;
        }
    }
    return t;
}

extern "C" x264cu_ctx *x264cu_open( int device, const x264cu_cfg *cfg )
{
    if( !cfg || cfg->width < 16 || cfg->height < 16 || cfg->bframes < 0 || cfg->bframes > X264CU_BFRAME_MAX || cfg->mv_range < 32 )
    { snprintf( g_err, sizeof(g_err), "x264cu_open: bad configuration" ); return NULL; }
    int ndev = x264cu_device_count();
    if( ndev <= 0 ) { snprintf( g_err, sizeof(g_err), "x264cu_open: no CUDA device (there is no CPU fallback)" ); return NULL; }
    if( device < 0 || device >= ndev ) { snprintf( g_err, sizeof(g_err), "x264cu_open: device %d of %d", device, ndev ); return NULL; }
    CUP( cudaSetDevice( device ) );
    x264cu_ctx *c = new x264cu_ctx();
    c->device = device;
    c->cfg = *cfg;
    if( c->cfg.search_passes < 0 ) c->cfg.search_passes = 2;      // 0 selects the exact raster-serial wavefront mode
    if( c->cfg.me_range < 4 ) c->cfg.me_range = 16;
    if( c->cfg.max_frames < 4 ) c->cfg.max_frames = 64;
    cudaDeviceProp prop;
    CUP( cudaGetDeviceProperties( &prop, device ) );
    c->sm_count = prop.multiProcessorCount;
    LaParams &P = c->P;
    P.mbw = ( cfg->width + 15 ) / 16; P.mbh = ( cfg->height + 15 ) / 16; P.nmb = P.mbw * P.mbh;
    P.lw = P.mbw * 8; P.lh = P.mbh * 8;
    // plane pitch: 1024 or 2048 whenever the padded width fits, so that the search/decision kernels run their
    // compile-time-pitch instantiations (row addresses become immediates); wider frames use a run-time pitch
    {
        const int padded = P.lw + 2*X264CU_PAD;
        const char *e = getenv( "X264CU_RUNTIME_PITCH" );      // testing aid: forces the run-time-pitch kernels
        P.stride = ( e && atoi( e ) ) || padded > X264CU_PITCH_B ? (int)align_up( padded, 128 ) + ( e && atoi( e ) && ( padded == X264CU_PITCH_A || padded == X264CU_PITCH_B ) ? 128 : 0 )
                 : padded <= X264CU_PITCH_A ? X264CU_PITCH_A : X264CU_PITCH_B;
    }
    P.plane_size = (int64_t)P.stride * ( P.lh + 2*X264CU_PAD );
    P.mv_range2 = 2 * cfg->mv_range;
    P.me_range = c->cfg.me_range;
    P.b_satd = cfg->subme > 1;
    P.me_method = P.b_satd ? ( cfg->me_method >= 1 ? 1 : 0 ) : 0;      // slicetype.c:50-59
    P.do_edges = cfg->do_edges != 0;
    P.aq = cfg->aq != 0;
    P.weighted_bipred = cfg->weighted_bipred != 0;
    P.cost_mv_n = 8 * cfg->mv_range;
    {   // tuning/test knob: force one of the two lane mappings (default: chosen per launch)
        const char *e = getenv( "X264CU_LANES" );
        P.lanes_per_block = e ? atoi( e ) : 0;
    }
    {
        const std::vector<uint32_t> ts = build_warp_tiles( P.mbw, P.mbh, X264CU_S_TILE_W, X264CU_S_TILE_H ), tf = build_warp_tiles( P.mbw, P.mbh, X264CU_F_TILE_W, X264CU_F_TILE_H );
        uint32_t *d = NULL;
        CUP( cudaMalloc( &d, ( ts.size() + tf.size() ) * 4 ) );
        CUP( cudaMemcpy( d, ts.data(), ts.size() * 4, cudaMemcpyHostToDevice ) );
        CUP( cudaMemcpy( d + ts.size(), tf.data(), tf.size() * 4, cudaMemcpyHostToDevice ) );
        c->d_tiles = d;
        P.tiles_s = d; P.ntiles_s = (int)ts.size(); P.tiles_f = d + ts.size(); P.ntiles_f = (int)tf.size();
    }
    c->plane_alloc = 4 * (size_t)P.plane_size + 256;
    c->stat_pitch = (int)align_up( P.mbh + 4, 4 );
    CUP( cudaStreamCreateWithFlags( &c->st, cudaStreamNonBlocking ) );
    CUP( cudaStreamCreateWithFlags( &c->st_in, cudaStreamNonBlocking ) );
    CUP( cudaStreamCreateWithFlags( &c->st_out, cudaStreamNonBlocking ) );
    CUP( cudaEventCreateWithFlags( &c->ev_res, cudaEventDisableTiming ) );
    for( int i = 0; i < 2; i++ ) CUP( cudaEventCreateWithFlags( &c->ev_up[i], cudaEventDisableTiming ) );
    // mv cost table: lambda * (log2(|d|+1)*2 + 1.718) rounded, lambda = 1 at X264_LOOKAHEAD_QP (analyse.c:153-157,186-188)
    {
        std::vector<uint16_t> tab( P.cost_mv_n + 1 );
        for( int i = 0; i <= P.cost_mv_n; i++ )
        {
            if( cfg->cost_mv ) tab[i] = cfg->cost_mv[i];
            else
            {
                float l = i ? log2f( (float)( i+1 ) ) * 2.0f + 1.718f : 0.718f;
                int v = (int)( 1 * l + .5f );
                tab[i] = (uint16_t)( v < 65535 ? v : 65535 );
            }
        }
        CUP( cudaMalloc( &c->d_cost_mv, tab.size() * sizeof(uint16_t) ) );
        CUP( cudaMemcpy( c->d_cost_mv, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice ) );
        P.cost_mv = c->d_cost_mv;
    }
    {   // x264_log2_lut (tables.c:66-85) is log2(1 + i/128) printed with 5 decimals
        float lut[128];
        for( int i = 0; i < 128; i++ ) lut[i] = (float)( round( log2( 1.0 + i / 128.0 ) * 1e5 ) / 1e5 );
        CUP( cudaMalloc( &c->d_log2_lut, sizeof(lut) ) );
        CUP( cudaMemcpy( c->d_log2_lut, lut, sizeof(lut), cudaMemcpyHostToDevice ) );
    }
    {   // x264_exp2_lut (tables.c:58-64) is round( 256 * ( 2^(i/64) - 1 ) )
        uint8_t lut[64];
        for( int i = 0; i < 64; i++ ) lut[i] = (uint8_t)lround( 256.0 * ( pow( 2.0, i / 64.0 ) - 1.0 ) );
        CUP( cudaMalloc( &c->d_exp2_lut, sizeof(lut) ) );
        CUP( cudaMemcpy( c->d_exp2_lut, lut, sizeof(lut), cudaMemcpyHostToDevice ) );
    }
    {
        const char *e = getenv( "X264CU_EXACT_WAVES" );
        c->exact_waves = e && atoi( e );
        c->serial_ctas_per_sm = search_serial_max_ctas_per_sm();
        if( c->serial_ctas_per_sm < 1 ) c->serial_ctas_per_sm = 1;
        c->serial_error = 0;
    }
    c->chroma_pitch = (int)align_up( cfg->width, 256 );
    c->luma_pitch = (int)align_up( cfg->width, 256 );
    for( int i = 0; i < 2; i++ )
    {
        CUP( cudaMalloc( &c->d_luma[i], (size_t)c->luma_pitch * cfg->height + 256 ) );
        c->d_chroma[i] = NULL;                    // allocated by the first x264cu_frame_upload_aq
        CUP( cudaEventCreateWithFlags( &c->luma_free[i], cudaEventDisableTiming ) );
    }
    c->pin_size = (size_t)64 << 20;            // cf. PAGE_LOCKED_BUF_SIZE 32 MiB (opencl.h:715)
    // one region (pin_size / NSLOT) holds everything one upload stages: luma + NV12 chroma + the per-MB tables
    while( c->pin_size / NSLOT < (size_t)c->luma_pitch * cfg->height * 3 / 2 + (size_t)P.nmb * 8 + ( 1 << 20 ) ) c->pin_size <<= 1;
    CUP( cudaHostAlloc( &c->pin, c->pin_size, cudaHostAllocDefault ) );
    c->jobs_size = (size_t)1 << 20;
    CUP( cudaHostAlloc( &c->h_jobs, c->jobs_size, cudaHostAllocDefault ) );
    CUP( cudaMalloc( &c->d_jobs, c->jobs_size ) );
    {   // per region: MV ping-pong fields of 2B+1 searches, a few weighted planes, the result gather block
        size_t region = (size_t)P.nmb * 4 * 2 * 40 + 4 * (size_t)P.plane_size + (size_t)P.nmb * 2 * 200;
        c->arena_size = NSLOT * align_up( region, 1 << 20 );
    }
    CUP( cudaMalloc( &c->arena, c->arena_size ) );
    for( int i = 0; i < NSLOT; i++ )
    {
        CUP( cudaEventCreateWithFlags( &c->slot_done[i], cudaEventDisableTiming ) );
        c->slot_pending[i] = false; c->slot_seq[i] = 0;
        c->pin_used[i] = c->jobs_used[i] = c->arena_used[i] = 0;
    }
    c->seq = 0; c->slot = 0; c->copies_head = 0;
    CUP( cudaEventCreate( &c->tm_a ) );
    CUP( cudaEventCreate( &c->tm_b ) );
    g_err[0] = 0;
    return c;
}

extern "C" int x264cu_info( x264cu_ctx *c, int *out )
{
    if( !c ) FAIL( "null context" );
    out[0] = c->P.mbw; out[1] = c->P.mbh; out[2] = c->P.lw; out[3] = c->P.lh; out[4] = c->P.stride; out[5] = c->sm_count;
    return 0;
}

extern "C" void x264cu_frame_free( x264cu_ctx *c, x264cu_frame *f )
{
    if( !c || !f ) return;
    cudaSetDevice( c->device );
    ctx_sync( c );
    cudaFree( f->planes ); cudaFree( f->inv_qscale ); cudaFree( f->costs );
    cudaFree( f->mvs ); cudaFree( f->mvcosts ); cudaFree( f->stats );
    cudaFree( f->prop ); cudaFree( f->qp_aq ); cudaFree( f->qp ); cudaFree( f->pixstats ); cudaFree( f->wres );
    for( size_t i = 0; i < c->frames.size(); i++ )
        if( c->frames[i] == f ) { c->frames.erase( c->frames.begin() + i ); break; }
    delete f;
}

extern "C" void x264cu_close( x264cu_ctx *c )
{
    if( !c ) return;
    cudaSetDevice( c->device );
    ctx_sync( c );
    while( !c->frames.empty() ) x264cu_frame_free( c, c->frames.back() );
    for( cudaEvent_t e : c->ev_pool ) cudaEventDestroy( e );
    for( int i = 0; i < 2; i++ ) { cudaFree( c->d_luma[i] ); cudaEventDestroy( c->luma_free[i] ); }
    cudaFree( c->d_pow_x ); cudaFree( c->d_pow_v );
    cudaFree( c->d_exp2_lut ); cudaFree( c->d_tiles ); cudaFree( c->d_chroma[0] ); cudaFree( c->d_chroma[1] );
    cudaFree( c->d_cost_mv ); cudaFree( c->d_log2_lut ); cudaFree( c->d_jobs ); cudaFree( c->arena );
    cudaFreeHost( c->pin ); cudaFreeHost( c->h_jobs );
    cudaEventDestroy( c->tm_a ); cudaEventDestroy( c->tm_b );
    for( int i = 0; i < NSLOT; i++ ) cudaEventDestroy( c->slot_done[i] );
    cudaEventDestroy( c->ev_res ); cudaEventDestroy( c->ev_up[0] ); cudaEventDestroy( c->ev_up[1] );
    cudaStreamDestroy( c->st_in ); cudaStreamDestroy( c->st_out ); cudaStreamDestroy( c->st );
    delete c;
}

extern "C" x264cu_frame *x264cu_frame_new( x264cu_ctx *c )
{
    if( !c ) { snprintf( g_err, sizeof(g_err), "null context" ); return NULL; }
    CUP( cudaSetDevice( c->device ) );
    const LaParams &P = c->P;
    const int B = c->cfg.bframes;
    x264cu_frame *f = new x264cu_frame();
    memset( f, 0, sizeof(*f) );
    size_t ntri = (size_t)( B+2 ) * ( B+2 ), nfield = 2 * (size_t)( B+1 );
    if( cudaMalloc( &f->planes, c->plane_alloc ) != cudaSuccess ||
        cudaMalloc( &f->inv_qscale, (size_t)P.nmb * 2 + 4 ) != cudaSuccess ||
        cudaMalloc( &f->costs, ntri * P.nmb * 2 ) != cudaSuccess ||
        cudaMalloc( &f->mvs, nfield * P.nmb * 4 ) != cudaSuccess ||
        cudaMalloc( &f->mvcosts, nfield * P.nmb * 4 ) != cudaSuccess ||
        cudaMalloc( &f->stats, ntri * c->stat_pitch * 4 ) != cudaSuccess ||
        cudaMalloc( &f->prop, (size_t)P.nmb * 4 ) != cudaSuccess ||
        cudaMalloc( &f->qp_aq, (size_t)P.nmb * 4 ) != cudaSuccess || cudaMalloc( &f->qp, (size_t)P.nmb * 4 ) != cudaSuccess ||
        cudaMalloc( &f->pixstats, 8 * sizeof(unsigned long long) ) != cudaSuccess || cudaMalloc( &f->wres, (size_t)( B+1 ) * 8 * 4 ) != cudaSuccess )
    {
        snprintf( g_err, sizeof(g_err), "x264cu_frame_new: out of device memory" );
        cudaGetLastError();
        cudaFree( f->planes ); cudaFree( f->inv_qscale ); cudaFree( f->costs ); cudaFree( f->mvs ); cudaFree( f->mvcosts ); cudaFree( f->stats );
        cudaFree( f->prop ); cudaFree( f->qp_aq ); cudaFree( f->qp ); cudaFree( f->pixstats ); cudaFree( f->wres );
        delete f;
        return NULL;
    }
    cudaMemset( f->prop, 0, (size_t)P.nmb * 4 );
    c->frames.push_back( f );
    return f;
}

// ---- timing spans ----
static cudaEvent_t get_event( x264cu_ctx *c )
{
    if( !c->ev_pool.empty() ) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate( &e ); return e;
}
struct Span
{
    x264cu_ctx *c; TimedSpan s; bool on;
    Span( x264cu_ctx *c_, int cls, int nlaunch ) : c( c_ ), on( c_->profiling )
    {
        c->launches += nlaunch; c->class_n[cls] += nlaunch;
        if( on ) { s.cls = cls; s.a = get_event( c ); s.b = get_event( c ); cudaEventRecord( s.a, c->st ); }
    }
    ~Span() { if( on ) { cudaEventRecord( s.b, c->st ); c->spans.push_back( s ); } }
};

// i_cost_est semantics completed on the host when the deferred copies are drained (slicetype.c:946-989).
// p: [0..2] raw sums of the evaluation (cost_est, cost_est_aq, intra_mbs), [4..5] intra sums of the frame
static void finish_stats( int32_t *dst, const uint8_t *p, int kind, int bias )
{
    const int32_t *raw = (const int32_t *)p, *intra = (const int32_t *)( p + 16 );
    if( kind == 0 )
    {   // I evaluation: the sums are the intra sums (slicetype.c:949-953,965-966)
        dst[0] = intra[0]; dst[1] = intra[1]; dst[2] = raw[2]; dst[3] = intra[0];
    }
    else
    {
        dst[3] = raw[0];
        dst[0] = kind == 2 ? (int32_t)( (uint64_t)raw[0] * 100 / ( 120 + bias ) ) : raw[0];     // :984-985
        dst[1] = raw[1]; dst[2] = raw[2];
    }
}

static void drain_copies( x264cu_ctx *c, uint64_t upto )
{   // the deferred host-side completion of x264_opencl_flush (slicetype-cl.c:64-69), in enqueue order
    while( c->copies_head < c->copies.size() && c->copies[c->copies_head].seq <= upto )
    {
        const PendingCopy &p = c->copies[c->copies_head++];
        if( p.kind < 0 ) memcpy( p.dst, p.src, p.bytes );
        else if( p.kind == 4 ) { if( *(const int *)p.src ) *(int *)p.dst = 1; }      // error flag of the serial search kernel
        else if( p.kind == 3 )
        {   // i_pixel_sum / i_pixel_ssd with the mean removed (ratecontrol.c:406-414); bias = mbw | mbh << 16
            const uint64_t *raw = (const uint64_t *)p.src; uint64_t *dst = (uint64_t *)p.dst;
            for( int i = 0; i < 3; i++ )
            {
                const uint64_t w = (uint64_t)( 16 * ( p.bias & 0xffff ) ) >> !!i, h = (uint64_t)( 16 * ( p.bias >> 16 ) ) >> !!i;
                dst[2*i] = raw[2*i];
                dst[2*i+1] = raw[2*i+1] - ( raw[2*i] * raw[2*i] + w * h / 2 ) / ( w * h );
            }
        }
        else finish_stats( (int32_t *)p.dst, (const uint8_t *)p.src, p.kind, p.bias );
    }
    if( c->copies_head == c->copies.size() ) { c->copies.clear(); c->copies_head = 0; }
}

static int wait_slot( x264cu_ctx *c, int i )
{
    if( !c->slot_pending[i] ) return 0;
    CU( cudaEventSynchronize( c->slot_done[i] ) );
    drain_copies( c, c->slot_seq[i] );
    c->slot_pending[i] = false;
    c->pin_used[i] = c->jobs_used[i] = c->arena_used[i] = 0;
    return 0;
}

// wait for every batch with sequence number <= seq, oldest first
static int wait_upto( x264cu_ctx *c, uint64_t seq )
{
    for( ;; )
    {
        int best = -1;
        for( int i = 0; i < NSLOT; i++ )
            if( c->slot_pending[i] && c->slot_seq[i] <= seq && ( best < 0 || c->slot_seq[i] < c->slot_seq[best] ) ) best = i;
        if( best < 0 ) return 0;
        if( wait_slot( c, best ) ) return -1;
    }
}

static int begin_batch( x264cu_ctx *c )
{
    c->seq++;
    c->batch_out = false;
    c->slot = (int)( c->seq % NSLOT );
    if( wait_upto( c, c->slot_pending[c->slot] ? c->slot_seq[c->slot] : 0 ) ) return -1;
    c->slot_seq[c->slot] = c->seq;
    return 0;
}

static int end_batch( x264cu_ctx *c, cudaStream_t last = NULL )
{
    CU( cudaEventRecord( c->slot_done[c->slot], last ? last : c->st ) );
    c->slot_pending[c->slot] = true;
    return 0;
}

static int ctx_sync( x264cu_ctx *c )
{
    CU( cudaStreamSynchronize( c->st_in ) );
    CU( cudaStreamSynchronize( c->st ) );
    CU( cudaStreamSynchronize( c->st_out ) );
    drain_copies( c, ~0ULL );
    for( int i = 0; i < NSLOT; i++ ) { c->slot_pending[i] = false; c->pin_used[i] = c->jobs_used[i] = c->arena_used[i] = 0; }
    for( TimedSpan &s : c->spans )
    {
        float ms = 0;
        cudaEventElapsedTime( &ms, s.a, s.b );
        c->class_ms[s.cls] += ms;
        c->ev_pool.push_back( s.a ); c->ev_pool.push_back( s.b );
    }
    c->spans.clear();
    return 0;
}

// Every enqueueing entry point runs inside one Batch.  A batch that is left without end() -- any `return -1` after
// work was enqueued -- is closed on the failure path: the context is drained (so that the staging, job and workspace
// regions it used are idle before they are handed out again) and the cache flags it set (searched / evaluated /
// analysed: "this field exists on the device") are rolled back, because their launches may never have been enqueued.
struct Batch
{
    x264cu_ctx *c; bool open, failed;
    explicit Batch( x264cu_ctx *c_ ) : c( c_ ), open( false ), failed( false )
    {
        c->batch_flags.clear();
        if( begin_batch( c ) ) failed = true; else open = true;
    }
    int end( cudaStream_t last = NULL ) { open = false; c->batch_flags.clear(); return end_batch( c, last ); }
    ~Batch()
    {
        if( !open ) return;
        char keep[sizeof(g_err)];
        memcpy( keep, g_err, sizeof(keep) );
        for( uint8_t *f : c->batch_flags ) *f = 0;
        c->batch_flags.clear();
        ctx_sync( c );
        cudaGetLastError();
        memcpy( g_err, keep, sizeof(keep) );
    }
};
static inline void mark( x264cu_ctx *c, uint8_t &flag ) { if( !flag ) { flag = 1; c->batch_flags.push_back( &flag ); } }

extern "C" int x264cu_flush( x264cu_ctx *c )
{
    if( !c ) FAIL( "null context" );
    CU( cudaSetDevice( c->device ) );
    if( ctx_sync( c ) ) return -1;
    if( c->serial_error ) { c->serial_error = 0; FAIL( "exact-mode search kernel: a row waited too long for the row below (launch not fully resident?)" ); }
    CU( cudaGetLastError() );
    return 0;
}

extern "C" int64_t x264cu_launch_count( x264cu_ctx *c ) { return c ? c->launches : 0; }

extern "C" int x264cu_kernel_times( x264cu_ctx *c, float *ms4, int *n4 )
{
    if( !c ) FAIL( "null context" );
    if( !c->spans.empty() && ctx_sync( c ) ) return -1;
    for( int i = 0; i < K_NCLASS; i++ )
    {
        if( ms4 ) ms4[i] = c->class_ms[i];
        if( n4 ) n4[i] = c->class_n[i];
        c->class_ms[i] = 0; c->class_n[i] = 0;
    }
    return 0;
}

extern "C" int x264cu_set_profiling( x264cu_ctx *c, int on )
{
    if( !c ) FAIL( "null context" );
    c->profiling = on != 0;
    return 0;
}

// pinned staging of the current batch's region: returns NULL when full (caller syncs everything and retries)
static uint8_t *pin_alloc( x264cu_ctx *c, size_t bytes )
{
    const size_t region = c->pin_size / NSLOT;
    size_t a = align_up( c->pin_used[c->slot], 256 );
    if( a + bytes > region ) return NULL;
    c->pin_used[c->slot] = a + bytes;
    return c->pin + c->slot * region + a;
}
// NOTE for callers: the retry path synchronises the context and hands out the region from offset 0 again, so everything
// staged earlier in the same batch must already have been ENQUEUED (its copy is then complete) -- stage and enqueue one
// array at a time, never stage two arrays and enqueue afterwards.
static uint8_t *pin_alloc_sync( x264cu_ctx *c, size_t bytes )
{
    uint8_t *p = pin_alloc( c, bytes );
    if( !p ) { if( ctx_sync( c ) ) return NULL; p = pin_alloc( c, bytes ); }
    if( !p ) snprintf( g_err, sizeof(g_err), "staging buffer too small for %zu bytes", bytes );
    return p;
}
static bool is_pinned( const void *p )
{
    cudaPointerAttributes at;
    if( cudaPointerGetAttributes( &at, p ) != cudaSuccess ) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// queue a device->host result copy: straight into pinned user memory, else via staging + deferred memcpy
static int queue_d2h( x264cu_ctx *c, void *dst, const void *dsrc, size_t bytes )
{
    if( !dst ) return 0;
    if( is_pinned( dst ) ) { CU( cudaMemcpyAsync( dst, dsrc, bytes, cudaMemcpyDeviceToHost, c->st ) ); return 0; }
    uint8_t *p = pin_alloc_sync( c, bytes );
    if( !p ) return -1;
    CU( cudaMemcpyAsync( p, dsrc, bytes, cudaMemcpyDeviceToHost, c->st ) );
    c->copies.push_back( { dst, p, bytes, -1, 0, c->seq } );
    return 0;
}

static inline uint8_t *planeF( x264cu_ctx *c, x264cu_frame *f ) { return f->planes + (size_t)X264CU_PAD * c->P.stride + X264CU_PAD; }
static inline int tri_index( x264cu_ctx *c, int d0, int d1 ) { return d0 * ( c->cfg.bframes + 2 ) + d1; }
static inline int field_index( x264cu_ctx *c, int list, int dist ) { return list * ( c->cfg.bframes + 1 ) + dist - 1; }

// aq-mode 2 / 3 call powf( energy + 1, 0.125f ) (ratecontrol.c:361).  energy < 2^24, so the argument is an exact integer in
// float; the device computes sqrt(sqrt(sqrt((double)x))) rounded to float, and this table lists every x of the domain where the
// HOST's libm -- the one the reference encoder links -- returns something else (glibc 2.39: 0.07 % of the inputs; powf is
// not correctly rounded).  Built once per process by running powf over the whole domain (~0.5 s), on first use of those modes.
static int ensure_pow_table( x264cu_ctx *c )
{
    if( c->d_pow_x ) return 0;
    static std::vector<uint32_t> hx;
    static std::vector<float> hv;
    static bool built = false;
    if( !built )
    {
        for( uint32_t e = 0; e < ( 1u << 24 ); e++ )
        {
            volatile float x = (float)e * 1.f + 1;
            const float a = powf( x, 0.125f );
            const float b = (float)sqrt( sqrt( sqrt( (double)x ) ) );
            if( a != b ) { float xv = x; uint32_t bits; memcpy( &bits, &xv, 4 ); hx.push_back( bits ); hv.push_back( a ); }
        }
        built = true;       // positive floats: bit patterns ascend with the value, so hx is sorted
    }
    c->pow_n = (int)hx.size();
    CU( cudaMalloc( &c->d_pow_x, ( hx.size() + 1 ) * 4 ) );
    CU( cudaMalloc( &c->d_pow_v, ( hv.size() + 1 ) * 4 ) );
    if( c->pow_n )
    {
        CU( cudaMemcpy( c->d_pow_x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice ) );
        CU( cudaMemcpy( c->d_pow_v, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice ) );
    }
    return 0;
}

// h->param.rc.f_aq_strength * 1.0397f as one single-precision product (ratecontrol.c:371)
static float aq_strength_f32( float s ) { volatile float r = s * 1.0397f; return r; }

static void reset_caches( x264cu_frame *f )
{   // mc.c:473-481: i_cost_est = -1, MV sentinels = 0x7FFF
    memset( f->searched, 0, sizeof(f->searched) );
    memset( f->evaluated, 0, sizeof(f->evaluated) );
    memset( f->wanalysed, 0, sizeof(f->wanalysed) );
    f->any_eval = false;
}

// adaptive quantisation between the downscale and the intra costs (x264cu_frame_upload_aq*)
struct AqRun { const uint8_t *luma; int ls; const uint8_t *chroma; int cs; int mode; float strength; const float *quant_offsets; };

// downscale (+ the per-frame resets), [adaptive quantisation], intra costs: 2 launches per frame, 3-5 with AQ
static int upload_common( x264cu_ctx *c, x264cu_frame *f, const uint8_t *d_luma, int pitch, bool fill_inv_qscale, const AqRun *aq = NULL )
{
    const LaParams &P = c->P;
    const int B = c->cfg.bframes;
    reset_caches( f );
    const int ntri = ( B+2 ) * ( B+2 );
    {
        Span s( c, K_LOWRES, 1 );
        launch_lowres( d_luma, pitch, c->cfg.width, c->cfg.height, f->planes, P, f->stats, ntri * c->stat_pitch,
                       fill_inv_qscale ? f->inv_qscale : NULL, aq ? f->pixstats : NULL, c->st );
    }
    if( aq )
    {
        if( aq->mode >= 2 && ensure_pow_table( c ) ) return -1;
        // pixstats[6..7] (unused by the sums) are the three floats the auto-variance reduction hands to its apply kernel
        c->launches += launch_aq( aq->luma, aq->ls, aq->chroma, aq->cs, c->cfg.width, c->cfg.height, P, aq->mode,
                                  aq->mode == 1 ? aq_strength_f32( aq->strength ) : aq->strength, aq->quant_offsets,
                                  c->d_log2_lut, c->d_exp2_lut, f->qp_aq, f->inv_qscale, f->pixstats, c->d_pow_x, c->d_pow_v, c->pow_n,
                                  (float *)( f->pixstats + 6 ), c->st );
        f->have_aq = true;
    }
    {
        Span s( c, K_INTRA, 1 );
        const uint8_t *F = f->planes + (size_t)X264CU_PAD * P.stride + X264CU_PAD;
        launch_intra( F, f->inv_qscale, f->costs, f->stats, f->stats + P.mbh, P, c->st );
    }
    CU( cudaGetLastError() );
    f->uploaded = true;
    return 0;
}

extern "C" int x264cu_frame_upload( x264cu_ctx *c, x264cu_frame *f, const uint8_t *luma, int stride, const uint16_t *inv_qscale )
{
    if( !c || !f || !luma ) FAIL( "x264cu_frame_upload: null argument" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    const LaParams &P = c->P;
    const int W = c->cfg.width, H = c->cfg.height;
    int sel = c->luma_sel; c->luma_sel ^= 1;
    const uint8_t *src = luma; int spitch = stride;
    if( !is_pinned( luma ) )
    {   // pageable caller memory: stage through the page-locked buffer (slicetype-cl.c:166-168)
        uint8_t *p = pin_alloc_sync( c, (size_t)W * H );
        if( !p ) return -1;
        for( int y = 0; y < H; y++ ) memcpy( p + (size_t)y*W, luma + (size_t)y*stride, W );
        src = p; spitch = W;
    }
    // the luma crosses PCIe on its own stream, ahead of the compute queue: it only has to wait until the lowres
    // kernel that last read this staging buffer is done, and the compute queue only waits for this copy
    CU( cudaStreamWaitEvent( c->st_in, c->luma_free[sel], 0 ) );
    CU( cudaMemcpy2DAsync( c->d_luma[sel], c->luma_pitch, src, spitch, W, H, cudaMemcpyHostToDevice, c->st_in ) );
    CU( cudaEventRecord( c->ev_up[sel], c->st_in ) );
    CU( cudaStreamWaitEvent( c->st, c->ev_up[sel], 0 ) );
    if( inv_qscale )
    {
        const uint16_t *q = inv_qscale;
        if( !is_pinned( q ) )
        {
            uint8_t *p = pin_alloc_sync( c, (size_t)P.nmb * 2 );
            if( !p ) return -1;
            memcpy( p, inv_qscale, (size_t)P.nmb * 2 );
            q = (const uint16_t *)p;
        }
        CU( cudaMemcpyAsync( f->inv_qscale, q, (size_t)P.nmb * 2, cudaMemcpyHostToDevice, c->st ) );
    }
    if( upload_common( c, f, c->d_luma[sel], c->luma_pitch, inv_qscale == NULL ) ) return -1;
    CU( cudaEventRecord( c->luma_free[sel], c->st ) );
    return B.end();
}

// x264_adaptive_quant_frame + x264_frame_init_lowres (encoder.c:3417-3421), SURVEY.md section 8f row 3
extern "C" int x264cu_frame_upload_aq( x264cu_ctx *c, x264cu_frame *f, const uint8_t *luma, int stride, const uint8_t *chroma, int cstride,
                                       const x264cu_aq_cfg *aq, const float *quant_offsets, const x264cu_aq_out *out )
{
    if( !c || !f || !luma || !chroma || !aq ) FAIL( "x264cu_frame_upload_aq: null argument" );
    if( aq->mode < 0 || aq->mode > 3 ) FAIL( "x264cu_frame_upload_aq: aq-mode %d", aq->mode );
    if( ( c->cfg.width | c->cfg.height ) & 1 ) FAIL( "x264cu_frame_upload_aq: 4:2:0 needs even dimensions" );
    CU( cudaSetDevice( c->device ) );
    const LaParams &P = c->P;
    const int W = c->cfg.width, H = c->cfg.height;
    for( int i = 0; i < 2; i++ )
        if( !c->d_chroma[i] ) CU( cudaMalloc( &c->d_chroma[i], (size_t)c->chroma_pitch * ( H / 2 ) + 256 ) );
    Batch B( c );
    if( B.failed ) return -1;
    int sel = c->luma_sel; c->luma_sel ^= 1;
    const uint8_t *src = luma, *csrc = chroma; int spitch = stride, cpitch = cstride;
    CU( cudaStreamWaitEvent( c->st_in, c->luma_free[sel], 0 ) );
    // one plane at a time: stage, enqueue, then stage the next (see pin_alloc_sync)
    if( !is_pinned( luma ) )
    {
        uint8_t *p = pin_alloc_sync( c, (size_t)W * H );
        if( !p ) return -1;
        for( int y = 0; y < H; y++ ) memcpy( p + (size_t)y*W, luma + (size_t)y*stride, W );
        src = p; spitch = W;
    }
    CU( cudaMemcpy2DAsync( c->d_luma[sel], c->luma_pitch, src, spitch, W, H, cudaMemcpyHostToDevice, c->st_in ) );
    if( !is_pinned( chroma ) )
    {
        uint8_t *p = pin_alloc_sync( c, (size_t)W * ( H / 2 ) );
        if( !p ) return -1;
        for( int y = 0; y < H / 2; y++ ) memcpy( p + (size_t)y*W, chroma + (size_t)y*cstride, W );
        csrc = p; cpitch = W;
    }
    CU( cudaMemcpy2DAsync( c->d_chroma[sel], c->chroma_pitch, csrc, cpitch, W, H / 2, cudaMemcpyHostToDevice, c->st_in ) );
    CU( cudaEventRecord( c->ev_up[sel], c->st_in ) );
    CU( cudaStreamWaitEvent( c->st, c->ev_up[sel], 0 ) );
    const float *d_qo = NULL;
    if( quant_offsets && aq->mode )
    {   // f->qp doubles as the device staging of the caller's quant_offsets (x264_picture_t.prop.quant_offsets)
        const float *q = quant_offsets;
        if( !is_pinned( q ) )
        {
            uint8_t *p = pin_alloc_sync( c, (size_t)P.nmb * 4 );
            if( !p ) return -1;
            memcpy( p, quant_offsets, (size_t)P.nmb * 4 );
            q = (const float *)p;
        }
        CU( cudaMemcpyAsync( f->qp, q, (size_t)P.nmb * 4, cudaMemcpyHostToDevice, c->st ) );
        d_qo = f->qp;
    }
    const AqRun run = { c->d_luma[sel], c->luma_pitch, c->d_chroma[sel], c->chroma_pitch, aq->mode, aq->strength, d_qo };
    if( upload_common( c, f, c->d_luma[sel], c->luma_pitch, true, &run ) ) return -1;
    CU( cudaEventRecord( c->luma_free[sel], c->st ) );
    if( out )
    {
        if( queue_d2h( c, out->qp_offset, f->qp_aq, (size_t)P.nmb * 4 ) ) return -1;
        if( queue_d2h( c, out->inv_qscale, f->inv_qscale, (size_t)P.nmb * 2 ) ) return -1;
        if( out->pixel_stats )
        {
            uint8_t *p = pin_alloc_sync( c, 64 );
            if( !p ) return -1;
            CU( cudaMemcpyAsync( p, f->pixstats, 48, cudaMemcpyDeviceToHost, c->st ) );
            c->copies.push_back( { out->pixel_stats, p, 48, 3, P.mbw | ( P.mbh << 16 ), c->seq } );
        }
    }
    return B.end();
}

extern "C" int x264cu_frame_upload_aq_device( x264cu_ctx *c, x264cu_frame *f, const uint8_t *d_luma, int stride, const uint8_t *d_chroma, int cstride,
                                              const x264cu_aq_cfg *aq, const float *d_quant_offsets )
{
    if( !c || !f || !d_luma || !d_chroma || !aq ) FAIL( "x264cu_frame_upload_aq_device: null argument" );
    if( aq->mode < 0 || aq->mode > 3 ) FAIL( "x264cu_frame_upload_aq_device: aq-mode %d", aq->mode );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    const LaParams &P = c->P;
    const AqRun run = { d_luma, stride, d_chroma, cstride, aq->mode, aq->strength, aq->mode ? d_quant_offsets : NULL };
    if( upload_common( c, f, d_luma, stride, true, &run ) ) return -1;
    return B.end();
}

// for callers that ran x264_adaptive_quant_frame on the host (the unmodified encoder does, encoder.c:3417): hand the
// luma statistics x264cu_weights_analyse needs to the device.  sum / ssd are i_pixel_sum[0] / i_pixel_ssd[0] AFTER the
// mean removal (ratecontrol.c:406-414); the raw sum of squares is recovered exactly.
extern "C" int x264cu_frame_set_pixel_stats( x264cu_ctx *c, x264cu_frame *f, uint64_t sum, uint64_t ssd )
{
    if( !c || !f ) FAIL( "x264cu_frame_set_pixel_stats: null argument" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    const uint64_t wh = (uint64_t)( 16 * c->P.mbw ) * ( 16 * c->P.mbh );
    uint64_t *p = (uint64_t *)pin_alloc_sync( c, 64 );
    if( !p ) return -1;
    memset( p, 0, 64 );
    p[0] = sum; p[1] = ssd + ( sum * sum + wh / 2 ) / wh;
    CU( cudaMemcpyAsync( f->pixstats, p, 64, cudaMemcpyHostToDevice, c->st ) );
    f->have_aq = true;
    return B.end();
}

// x264_weights_analyse( h, fenc, fref, 1 ) (slicetype.c:284-501), SURVEY.md section 8f row 2
extern "C" int x264cu_weights_analyse( x264cu_ctx *c, x264cu_frame *fenc, x264cu_frame *fref, int dist, int numslices, x264cu_weight *out )
{
    if( !c || !fenc || !fref || dist < 1 || dist > c->cfg.bframes + 1 ) FAIL( "x264cu_weights_analyse: bad argument" );
    if( !fenc->uploaded || !fref->uploaded || !fenc->have_aq || !fref->have_aq )
        FAIL( "x264cu_weights_analyse: both frames need the pixel statistics of x264cu_frame_upload_aq" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    const LaParams &P = c->P;
    int *wres = fenc->wres + 8 * ( dist - 1 );
    // weight_cost_init_luma (:76-103): motion-compensate the reference when the list-0 field has been searched
    const int16_t *mvs = fenc->searched[0][dist-1] ? fenc->mvs + (size_t)field_index( c, 0, dist ) * P.nmb * 2 : NULL;
    // the intra costs exist since upload; the reference computes them here if need be (:361-366), which also makes a
    // later I evaluation a cache hit
    fenc->any_eval = true;
    mark( c, fenc->wanalysed[dist-1] );
    c->launches += 3;
    launch_weights_analyse( planeF( c, fenc ), planeF( c, fref ), mvs, fenc->costs, fenc->pixstats, fref->pixstats, wres,
                            numslices > 0 ? numslices : 1, P, c->st );
    CU( cudaGetLastError() );
    if( queue_d2h( c, out, wres, sizeof(x264cu_weight) ) ) return -1;
    return B.end();
}

extern "C" int x264cu_frame_upload_device( x264cu_ctx *c, x264cu_frame *f, const uint8_t *d_luma, int stride, const uint16_t *d_inv_qscale )
{
    if( !c || !f || !d_luma ) FAIL( "x264cu_frame_upload_device: null argument" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    if( d_inv_qscale )
        CU( cudaMemcpyAsync( f->inv_qscale, d_inv_qscale, (size_t)c->P.nmb * 2, cudaMemcpyDeviceToDevice, c->st ) );
    if( upload_common( c, f, d_luma, stride, d_inv_qscale == NULL ) ) return -1;
    return B.end();
}

extern "C" int x264cu_frame_get_lowres( x264cu_ctx *c, x264cu_frame *f, int plane, uint8_t *out )
{
    if( !c || !f || plane < 0 || plane > 3 ) FAIL( "x264cu_frame_get_lowres: bad argument" );
    CU( cudaSetDevice( c->device ) );
    const LaParams &P = c->P;
    if( ctx_sync( c ) ) return -1;
    CU( cudaMemcpy2D( out, P.lw + 2*X264CU_PAD, f->planes + plane*P.plane_size, P.stride, P.lw + 2*X264CU_PAD,
                      P.lh + 2*X264CU_PAD, cudaMemcpyDeviceToHost ) );
    return 0;
}

// ---- job tables and workspace ----
template<typename T> static int push_jobs( x264cu_ctx *c, const std::vector<T> &v, const T **d_out )
{
    const size_t region = c->jobs_size / NSLOT, bytes = v.size() * sizeof(T);
    size_t a = align_up( c->jobs_used[c->slot], 16 );
    if( a + bytes > region )
    {
        if( bytes > region ) FAIL( "job table too large" );
        if( ctx_sync( c ) ) return -1;
        a = 0;
    }
    uint8_t *h = c->h_jobs + c->slot * region + a, *d = c->d_jobs + c->slot * region + a;
    memcpy( h, v.data(), bytes );
    CU( cudaMemcpyAsync( d, h, bytes, cudaMemcpyHostToDevice, c->st ) );
    c->jobs_used[c->slot] = a + bytes;
    *d_out = (const T *)d;
    return 0;
}

// device workspace of the current batch's region; grows the arena (when idle) if one batch needs more
static uint8_t *arena_alloc( x264cu_ctx *c, size_t bytes )
{
    size_t region = c->arena_size / NSLOT;
    size_t a = align_up( c->arena_used[c->slot], 256 );
    if( a + bytes > region )
    {
        if( ctx_sync( c ) ) return NULL;
        a = 0;
        if( bytes > region )
        {
            cudaFree( c->arena );
            c->arena_size = NSLOT * align_up( bytes * 2, 1 << 20 );
            if( cudaMalloc( &c->arena, c->arena_size ) != cudaSuccess )
            { snprintf( g_err, sizeof(g_err), "workspace allocation of %zu bytes failed", c->arena_size ); c->arena = NULL; c->arena_size = 0; return NULL; }
            region = c->arena_size / NSLOT;
        }
    }
    c->arena_used[c->slot] = a + bytes;
    return c->arena + c->slot * region + a;
}

struct SearchTask { x264cu_frame *fenc, *fref; int list, dist; x264cu_weight w; };      // w.enabled == X264CU_WEIGHT_DEVICE: fenc->wres[dist-1]


// K passes of the dependency-free schedule over a set of (fenc, ref, list, dist) fields, each pass
// ONE launch over all fields (x264_opencl_motionsearch batched as in x264_opencl_slicetype_prep).
static int run_searches( x264cu_ctx *c, const std::vector<SearchTask> &tasks )
{
    const LaParams &P = c->P;
    const int K = c->cfg.search_passes;
    if( !tasks.empty() )
    {
        const size_t field_bytes = align_up( (size_t)P.nmb * 4, 256 ), plane_bytes = align_up( (size_t)P.plane_size, 256 );
        size_t need = tasks.size() * 2 * field_bytes;
        for( const SearchTask &t : tasks ) if( t.w.enabled ) need += plane_bytes;
        // ONE allocation per batch before anything is enqueued: a second one could recycle (or re-allocate) the region
        // while the pointers handed out by the first are still waiting to be used
        uint8_t *ws = arena_alloc( c, need + 256 );
        if( !ws ) return -1;
        int *derr = (int *)ws; ws += 256;
        std::vector<uint8_t *> tmp0( tasks.size() ), tmp1( tasks.size() ), wpl( tasks.size(), NULL );
        for( size_t t = 0; t < tasks.size(); t++ )
        {
            tmp0[t] = ws; ws += field_bytes; tmp1[t] = ws; ws += field_bytes;
            if( tasks[t].w.enabled )
            {
                wpl[t] = ws; ws += plane_bytes;
                c->launches++;
                if( tasks[t].w.enabled == X264CU_WEIGHT_DEVICE )
                    launch_weight_plane_dev( tasks[t].fref->planes, wpl[t], P.plane_size, tasks[t].fenc->wres + 8 * ( tasks[t].dist - 1 ), c->st );
                else
                    launch_weight_plane( tasks[t].fref->planes, wpl[t], P.plane_size, tasks[t].w.scale, tasks[t].w.denom, tasks[t].w.offset, c->st );
            }
        }
        if( K == 0 )
        {   // exact raster-serial mode: neighbours are read from the field being written, wave by wave
            std::vector<SearchJob> sj( tasks.size() );
            for( size_t t = 0; t < tasks.size(); t++ )
            {
                const SearchTask &T = tasks[t];
                int fi = field_index( c, T.list, T.dist );
                SearchJob &S = sj[t];
                S.fenc = planeF( c, T.fenc ); S.ref = planeF( c, T.fref );
                S.ref_w = wpl[t] ? wpl[t] + (size_t)X264CU_PAD * P.stride + X264CU_PAD : S.ref;
                S.mv_out = T.fenc->mvs + (size_t)fi * P.nmb * 2;
                S.mv_in = S.mv_out;
                S.cost_out = T.fenc->mvcosts + (size_t)fi * P.nmb;
                S.wscale = T.w.scale; S.wdenom = T.w.enabled ? T.w.denom : -1; S.woffset = T.w.offset;
                S.wdev = T.w.enabled == X264CU_WEIGHT_DEVICE ? T.fenc->wres + 8 * ( T.dist - 1 ) : NULL;
                CU( cudaMemsetAsync( S.mv_out, 0, (size_t)P.nmb * 4, c->st ) );
                CU( cudaMemsetAsync( S.cost_out, 0, (size_t)P.nmb * 4, c->st ) );
            }
            const SearchJob *d;
            if( push_jobs( c, sj, &d ) ) return -1;
            if( c->exact_waves )
            {   // one launch per anti-diagonal wave (the first implementation; kept as a cross-check, X264CU_EXACT_WAVES=1)
                const int nwaves = P.mbw + 2 * ( P.mbh - 1 );
                Span s( c, K_SEARCH, nwaves );
                launch_search_waves( d, (int)sj.size(), P, c->st );
            }
            else
            {   // one persistent launch: every MV word of the field is its own "ready" flag (search_kernel_serial)
                for( const SearchJob &S : sj ) CU( cudaMemsetAsync( S.mv_out, 0x7F, (size_t)P.nmb * 4, c->st ) );
                CU( cudaMemsetAsync( derr, 0, 4, c->st ) );
                Span s( c, K_SEARCH, 0 );
                int nl = launch_search_serial( d, (int)sj.size(), P, derr, c->sm_count * c->serial_ctas_per_sm, c->st );
                if( nl < 0 ) FAIL( "exact-mode search kernel: cooperative launch refused (one field row per warp does not fit the device)" );
                c->launches += nl; c->class_n[K_SEARCH] += nl;
                uint8_t *p = pin_alloc_sync( c, 4 );
                if( !p ) return -1;
                CU( cudaMemcpyAsync( p, derr, 4, cudaMemcpyDeviceToHost, c->st ) );
                c->copies.push_back( { &c->serial_error, p, 4, 4, 0, c->seq } );
            }
        }
        for( int pass = 0; pass < K; pass++ )
        {
            std::vector<SearchJob> sj( tasks.size() );
            for( size_t t = 0; t < tasks.size(); t++ )
            {
                const SearchTask &T = tasks[t];
                int fi = field_index( c, T.list, T.dist );
                SearchJob &S = sj[t];
                S.fenc = planeF( c, T.fenc ); S.ref = planeF( c, T.fref );
                S.ref_w = wpl[t] ? wpl[t] + (size_t)X264CU_PAD * P.stride + X264CU_PAD : S.ref;
                S.mv_in = pass == 0 ? NULL : (const int16_t *)( ( pass & 1 ) ? tmp0[t] : tmp1[t] );
                S.mv_out = pass == K-1 ? T.fenc->mvs + (size_t)fi * P.nmb * 2 : (int16_t *)( ( pass & 1 ) ? tmp1[t] : tmp0[t] );
                S.cost_out = T.fenc->mvcosts + (size_t)fi * P.nmb;
                S.wscale = T.w.scale; S.wdenom = T.w.enabled ? T.w.denom : -1; S.woffset = T.w.offset;
                S.wdev = T.w.enabled == X264CU_WEIGHT_DEVICE ? T.fenc->wres + 8 * ( T.dist - 1 ) : NULL;
            }
            const SearchJob *d;
            if( push_jobs( c, sj, &d ) ) return -1;
            Span s( c, K_SEARCH, 1 );
            bool anyw = false;
            for( const SearchTask &t : tasks ) anyw = anyw || t.w.enabled;
            launch_search( d, (int)sj.size(), anyw, pass == 0, P, c->st );
        }
    }
    return 0;
}

// Order the decision jobs of a batch so that evaluations of one frame that share their list-1 side -- the family
// (., p1, b) over all p0 of the B evaluations, or the I / P evaluations of a frame -- are consecutive, and cut every
// family into chunks of <= X264CU_F_CHUNK jobs (final.cu final_kernel_c: one thread keeps the source block and the
// list-1 prediction for a whole chunk; intra cost and AQ factor are read once).  Jobs are independent (integer atomics,
// distinct outputs), so any order is legal.  Long chunks first: they should not be the tail of the launch.
static bool same_family( const FinalJob &a, const FinalJob &b )
{
    if( a.b_bidir != b.b_bidir || a.fenc != b.fenc || a.intra != b.intra || a.inv_qscale != b.inv_qscale ) return false;
    return !a.b_bidir || ( a.ref1 == b.ref1 && a.mv[1] == b.mv[1] && a.mvcost[1] == b.mvcost[1] );
}
#ifndef X264CU_F_SORT
#define X264CU_F_SORT 1          // 0: tuning builds only (jobs stay in the caller's order)
#endif
static void group_finals( std::vector<FinalJob> &finals, std::vector<int2> &chunks )
{
    if( X264CU_F_SORT ) std::stable_sort( finals.begin(), finals.end(), []( const FinalJob &a, const FinalJob &b )
    {
        if( a.b_bidir != b.b_bidir ) return a.b_bidir > b.b_bidir;
        if( a.fenc != b.fenc ) return a.fenc < b.fenc;
        if( !a.b_bidir ) return false;                      // the I / P evaluations of a frame keep their order
        if( a.ref1 != b.ref1 ) return a.ref1 < b.ref1;
        return a.mv[1] < b.mv[1];
    } );
    const int n = (int)finals.size();
    for( int i = 0; i < n; )
    {
        int j = i + 1;
        while( j < n && j - i < X264CU_F_CHUNK && same_family( finals[i], finals[j] ) ) j++;
        chunks.push_back( make_int2( i, j - i ) );
        i = j;
    }
    if( X264CU_F_SORT ) std::stable_sort( chunks.begin(), chunks.end(), [&]( const int2 &a, const int2 &b )
    {   // bidir work dominates; among bidir chunks the longer one first
        const int wa = finals[a.x].b_bidir ? a.y : 0, wb = finals[b.x].b_bidir ? b.y : 0;
        return wa > wb;
    } );
}

static int run_batch( x264cu_ctx *c, const x264cu_job *jobs, int n )
{
    const LaParams &P = c->P;
    const int B = c->cfg.bframes;
    std::vector<SearchTask> tasks;
    std::vector<FinalJob> finals;
    // pass 1 over the jobs in semantic order: cache bookkeeping exactly as slicetype_frame_cost does it
    for( int i = 0; i < n; i++ )
    {
        const x264cu_job &j = jobs[i];
        if( !j.fenc || !j.fref0 || !j.fref1 || j.d0 < 0 || j.d1 < 0 || j.d0 + j.d1 > B + 1 )
            FAIL( "x264cu_frame_cost: bad job %d (d0=%d d1=%d bframes=%d)", i, j.d0, j.d1, B );
        if( !j.fenc->uploaded || !j.fref0->uploaded || !j.fref1->uploaded ) FAIL( "x264cu_frame_cost: frame not uploaded" );
        // An I evaluation after any other evaluation of the frame is a cache hit in the reference
        // (i_cost_est[0][0] was filled by the intra section of that evaluation, slicetype.c:848,949-953)
        if( j.d0 == 0 && j.d1 == 0 && j.fenc->any_eval ) j.fenc->evaluated[0][0] = 1;
        j.fenc->any_eval = true;
        if( j.fenc->evaluated[j.d0][j.d1] ) continue;                  // i_cost_est >= 0 (slicetype.c:848)
        mark( c, j.fenc->evaluated[j.d0][j.d1] );
        FinalJob F;
        memset( &F, 0, sizeof(F) );
        F.fenc = planeF( c, j.fenc ); F.ref0 = planeF( c, j.fref0 ); F.ref1 = planeF( c, j.fref1 );
        F.b_bidir = j.d1 > 0; F.b_intra_only = j.d0 == 0 && j.d1 == 0;
        for( int l = 0; l < 2; l++ )
        {
            int dist = l ? j.d1 : j.d0;
            if( dist == 0 ) continue;
            int fi = field_index( c, l, dist );
            if( !j.fenc->searched[l][dist-1] )
            {   // do_search[l] (slicetype.c:855-867)
                mark( c, j.fenc->searched[l][dist-1] );
                SearchTask t = { j.fenc, l ? j.fref1 : j.fref0, l, dist, { 0, 0, 0, 0 } };
                if( l == 0 && j.d1 == 0 && j.weight.enabled )                      // w only when b == p1 (:859-864)
                {
                    if( j.weight.enabled == X264CU_WEIGHT_DEVICE && !j.fenc->wanalysed[dist-1] )
                        FAIL( "x264cu_frame_cost: X264CU_WEIGHT_DEVICE without a preceding x264cu_weights_analyse for distance %d", dist );
                    t.w = j.weight;
                }
                tasks.push_back( t );
            }
            F.mv[l] = j.fenc->mvs + (size_t)fi * P.nmb * 2;
            F.mvcost[l] = j.fenc->mvcosts + (size_t)fi * P.nmb;
        }
        if( F.b_bidir && j.fref1->searched[0][j.d0 + j.d1 - 1] )      // slicetype.c:629
            F.dmv_src = j.fref1->mvs + (size_t)field_index( c, 0, j.d0 + j.d1 ) * P.nmb * 2;
        F.intra = j.fenc->costs;
        F.inv_qscale = j.fenc->inv_qscale;
        int ti = tri_index( c, j.d0, j.d1 );
        F.lowres_costs = j.fenc->costs + (size_t)ti * P.nmb;
        F.row_satds = j.fenc->stats + (size_t)ti * c->stat_pitch;
        F.stats = F.row_satds + P.mbh;
        if( F.b_intra_only ) { F.row_satds = NULL; }                   // intra sums were produced at upload
        F.dist_scale_factor = ( j.d0 + j.d1 ) ? ( ( j.d0 << 8 ) + ( ( j.d0 + j.d1 ) >> 1 ) ) / ( j.d0 + j.d1 ) : 128;   // :869-870
        F.bipred_weight = P.weighted_bipred ? 64 - ( F.dist_scale_factor >> 2 ) : 32;                                    // :529
        finals.push_back( F );
    }
    if( run_searches( c, tasks ) ) return -1;
    if( !finals.empty() )
    {
        std::vector<int2> chunks;
        group_finals( finals, chunks );
        const FinalJob *d; const int2 *dch;
        if( push_jobs( c, finals, &d ) || push_jobs( c, chunks, &dch ) ) return -1;
        Span s( c, K_FINAL, 1 );
        launch_final( d, (int)finals.size(), dch, (int)chunks.size(), P, c->st );
    }
    CU( cudaGetLastError() );
    // result copies (slicetype-cl.c:254-280,513-535,613-649): every requested array is gathered on the
    // device into one staging block (one launch), which then crosses PCIe as ONE transfer -- straight
    // into the caller's memory when that is page-locked and laid out like the block, else through the
    // page-locked buffer with the per-array memcpy deferred to flush, as the reference does.
    struct HostOut { void *dst; size_t off, bytes; int kind; };
    std::vector<GatherSeg> segs;
    std::vector<HostOut> arrs, sts;
    size_t offA = 0, offS = 0;
    auto addA = [&]( void *dst, const void *dsrc, size_t bytes )
    {
        if( !dst ) return;
        segs.push_back( { (const uint8_t *)dsrc, (unsigned long long)offA, (unsigned)bytes } );
        arrs.push_back( { dst, offA, bytes, -1 } );
        offA += align_up( bytes, 64 );
    };
    for( int i = 0; i < n; i++ )
    {
        const x264cu_job &j = jobs[i];
        const x264cu_out &o = j.out;
        int ti = tri_index( c, j.d0, j.d1 );
        const int32_t *st = j.fenc->stats + (size_t)ti * c->stat_pitch;
        // order: intra_costs, lowres_costs, row_satds, mvs[0], mv_costs[0], mvs[1], mv_costs[1], intra_row_satds, intra_stats
        addA( o.intra_costs, j.fenc->costs, (size_t)P.nmb * 2 );
        addA( o.lowres_costs, j.fenc->costs + (size_t)ti * P.nmb, (size_t)P.nmb * 2 );
        addA( o.row_satds, j.d0 + j.d1 ? st : j.fenc->stats, (size_t)P.mbh * 4 );
        for( int l = 0; l < 2; l++ )
        {
            int dist = l ? j.d1 : j.d0;
            if( dist == 0 ) continue;
            int fi = field_index( c, l, dist );
            addA( o.mvs[l], j.fenc->mvs + (size_t)fi * P.nmb * 2, (size_t)P.nmb * 4 );
            addA( o.mv_costs[l], j.fenc->mvcosts + (size_t)fi * P.nmb, (size_t)P.nmb * 4 );
        }
        addA( o.intra_row_satds, j.fenc->stats, (size_t)P.mbh * 4 );
        addA( o.intra_stats, j.fenc->stats + P.mbh, 8 );
    }
    for( int i = 0; i < n; i++ )
    {
        const x264cu_job &j = jobs[i];
        if( !j.out.stats ) continue;
        sts.push_back( { j.out.stats, offS, 32, ( j.d0 + j.d1 == 0 ) ? 0 : ( j.d1 > 0 ? 2 : 1 ) } );
        offS += 32;
    }
    if( segs.empty() && sts.empty() ) return 0;
    const size_t baseS = align_up( offA, 256 ), total = baseS + offS;
    uint8_t *dstage = arena_alloc( c, total );
    if( !dstage ) return -1;
    {
        size_t k = 0;
        for( int i = 0; i < n; i++ )
        {
            const x264cu_job &j = jobs[i];
            if( !j.out.stats ) continue;
            int ti = tri_index( c, j.d0, j.d1 );
            const int32_t *st = j.fenc->stats + (size_t)ti * c->stat_pitch;
            segs.push_back( { (const uint8_t *)( st + P.mbh ), (unsigned long long)( baseS + sts[k].off ), 12u } );
            segs.push_back( { (const uint8_t *)( j.fenc->stats + P.mbh ), (unsigned long long)( baseS + sts[k].off + 16 ), 8u } );
            k++;
        }
    }
    const GatherSeg *dsegs;
    if( push_jobs( c, segs, &dsegs ) ) return -1;
    c->launches++;
    launch_gather( dsegs, (int)segs.size(), dstage, c->st );
    CU( cudaGetLastError() );
    // the results cross PCIe behind the compute queue, on their own stream, so the next batch's kernels overlap them
    CU( cudaEventRecord( c->ev_res, c->st ) );
    CU( cudaStreamWaitEvent( c->st_out, c->ev_res, 0 ) );
    c->batch_out = true;
    if( offA )
    {
        bool direct = is_pinned( arrs[0].dst );
        uint8_t *base = (uint8_t *)arrs[0].dst - arrs[0].off;
        for( const HostOut &h : arrs ) direct = direct && (uint8_t *)h.dst == base + h.off;
        if( direct )
            CU( cudaMemcpyAsync( base, dstage, offA, cudaMemcpyDeviceToHost, c->st_out ) );
        else
        {
            uint8_t *p = pin_alloc( c, offA );
            if( !p )
            {   // staging full: drain what is queued; the kernels above are then complete as well
                if( ctx_sync( c ) ) return -1;
                p = pin_alloc( c, offA );
                if( !p ) FAIL( "staging buffer too small for %zu result bytes", offA );
                // the arena was recycled by the sync but its contents are intact until the next batch
            }
            CU( cudaMemcpyAsync( p, dstage, offA, cudaMemcpyDeviceToHost, c->st_out ) );
            for( const HostOut &h : arrs ) c->copies.push_back( { h.dst, p + h.off, h.bytes, -1, 0, c->seq } );
        }
    }
    if( offS )
    {
        uint8_t *p = pin_alloc( c, offS );
        if( !p )
        {
            if( ctx_sync( c ) ) return -1;
            p = pin_alloc( c, offS );
            if( !p ) FAIL( "staging buffer too small" );
        }
        CU( cudaMemcpyAsync( p, dstage + baseS, offS, cudaMemcpyDeviceToHost, c->st_out ) );
        for( const HostOut &h : sts ) c->copies.push_back( { h.dst, p + h.off, 32, h.kind, c->cfg.bframe_bias, c->seq } );
    }
    return 0;
}

extern "C" int x264cu_frame_cost_batch( x264cu_ctx *c, const x264cu_job *jobs, int n )
{
    if( !c || ( !jobs && n ) ) FAIL( "x264cu_frame_cost_batch: null argument" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    if( run_batch( c, jobs, n ) ) return -1;
    return B.end( c->batch_out ? c->st_out : c->st );
}

extern "C" int x264cu_motionsearch( x264cu_ctx *c, x264cu_frame *fenc, x264cu_frame *fref, int list, int dist,
                                    const x264cu_weight *w, int16_t *mvs, int32_t *mv_costs )
{
    if( !c || !fenc || !fref || list < 0 || list > 1 || dist < 1 || dist > c->cfg.bframes + 1 ) FAIL( "x264cu_motionsearch: bad argument" );
    if( !fenc->uploaded || !fref->uploaded ) FAIL( "x264cu_motionsearch: frame not uploaded" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    if( !fenc->searched[list][dist-1] )
    {
        mark( c, fenc->searched[list][dist-1] );
        std::vector<SearchTask> tasks( 1 );
        tasks[0] = { fenc, fref, list, dist, { 0, 0, 0, 0 } };
        if( list == 0 && w && w->enabled )
        {
            if( w->enabled == X264CU_WEIGHT_DEVICE && !fenc->wanalysed[dist-1] )
                FAIL( "x264cu_motionsearch: X264CU_WEIGHT_DEVICE without a preceding x264cu_weights_analyse for distance %d", dist );
            tasks[0].w = *w;
        }
        if( run_searches( c, tasks ) ) return -1;
        CU( cudaGetLastError() );
    }
    int fi = field_index( c, list, dist );
    if( queue_d2h( c, mvs, fenc->mvs + (size_t)fi * c->P.nmb * 2, (size_t)c->P.nmb * 4 ) ) return -1;
    if( queue_d2h( c, mv_costs, fenc->mvcosts + (size_t)fi * c->P.nmb, (size_t)c->P.nmb * 4 ) ) return -1;
    return B.end();
}

extern "C" int x264cu_search_batch( x264cu_ctx *c, const x264cu_search_job *jobs, int n )
{
    if( !c || ( n > 0 && !jobs ) ) FAIL( "x264cu_search_batch: null argument" );
    CU( cudaSetDevice( c->device ) );
    Batch B( c );
    if( B.failed ) return -1;
    std::vector<SearchTask> tasks;
    for( int i = 0; i < n; i++ )
    {
        const x264cu_search_job &j = jobs[i];
        if( !j.fenc || !j.fref || j.list < 0 || j.list > 1 || j.dist < 1 || j.dist > c->cfg.bframes + 1 ) FAIL( "x264cu_search_batch: bad job %d", i );
        if( !j.fenc->uploaded || !j.fref->uploaded ) FAIL( "x264cu_search_batch: frame not uploaded" );
        if( j.fenc->searched[j.list][j.dist-1] ) continue;
        mark( c, j.fenc->searched[j.list][j.dist-1] );
        SearchTask t = { j.fenc, j.fref, j.list, j.dist, { 0, 0, 0, 0 } };
        if( j.list == 0 && j.weight.enabled )
        {
            if( j.weight.enabled == X264CU_WEIGHT_DEVICE && !j.fenc->wanalysed[j.dist-1] )
                FAIL( "x264cu_search_batch: X264CU_WEIGHT_DEVICE without a preceding x264cu_weights_analyse for distance %d", j.dist );
            t.w = j.weight;
        }
        tasks.push_back( t );
    }
    if( run_searches( c, tasks ) ) return -1;
    CU( cudaGetLastError() );
    return B.end();
}

extern "C" int x264cu_field_device( x264cu_ctx *c, x264cu_frame *f, int list, int dist, void **d_mvs, void **d_mv_costs )
{
    if( !c || !f || list < 0 || list > 1 || dist < 1 || dist > c->cfg.bframes + 1 ) FAIL( "x264cu_field_device: bad argument" );
    const int fi = field_index( c, list, dist );
    if( d_mvs ) *d_mvs = f->mvs + (size_t)fi * c->P.nmb * 2;
    if( d_mv_costs ) *d_mv_costs = f->mvcosts + (size_t)fi * c->P.nmb;
    return 0;
}

extern "C" int x264cu_field_set_searched( x264cu_ctx *c, x264cu_frame *f, int list, int dist )
{
    if( !c || !f || list < 0 || list > 1 || dist < 1 || dist > c->cfg.bframes + 1 ) FAIL( "x264cu_field_set_searched: bad argument" );
    f->searched[list][dist-1] = 1;
    return 0;
}

// sequence number of the most recently enqueued batch / wait for everything up to a sequence number:
// lets a caller keep NSLOT-1 batches in flight and still pick up results in order (x264cu_stream_sync)
extern "C" uint64_t x264cu_last_batch( x264cu_ctx *c ) { return c ? c->seq : 0; }
extern "C" uint64_t x264cu_fence( x264cu_ctx *c )
{
    if( !c || cudaSetDevice( c->device ) != cudaSuccess ) return 0;
    Batch B( c );
    if( B.failed || B.end() ) return 0;
    return c->seq;
}
extern "C" int x264cu_wait( x264cu_ctx *c, uint64_t seq )
{
    if( !c ) FAIL( "null context" );
    CU( cudaSetDevice( c->device ) );
    if( wait_upto( c, seq ) ) return -1;
    if( c->serial_error ) { c->serial_error = 0; FAIL( "exact-mode search kernel: a row waited too long for the row below (launch not fully resident?)" ); }
    return 0;
}

extern "C" int x264cu_frame_cost( x264cu_ctx *c, const x264cu_job *job ) { return x264cu_frame_cost_batch( c, job, 1 ); }

// ------------------------------------------------------------------------------------------------
// MB-tree: macroblock_tree_propagate / macroblock_tree_finish on the device
// ------------------------------------------------------------------------------------------------
extern "C" int x264cu_mbtree_reset( x264cu_ctx *c, x264cu_frame *f )
{   // memset( i_propagate_cost, 0, ... ) (slicetype.c:1128,1139,1145)
    if( !c || !f ) FAIL( "x264cu_mbtree_reset: null argument" );
    CU( cudaSetDevice( c->device ) );
    CU( cudaMemsetAsync( f->prop, 0, (size_t)c->P.nmb * 4, c->st ) );
    return 0;
}

extern "C" int x264cu_mbtree_swap( x264cu_ctx *c, x264cu_frame *a, x264cu_frame *b )
{   // XCHG( uint16_t*, frames[last_nonb]->i_propagate_cost, frames[0]->i_propagate_cost ) (slicetype.c:1127,1180)
    if( !c || !a || !b ) FAIL( "x264cu_mbtree_swap: null argument" );
    uint32_t *t = a->prop; a->prop = b->prop; b->prop = t;
    return 0;
}

extern "C" int x264cu_mbtree_propagate( x264cu_ctx *c, x264cu_frame *fb, x264cu_frame *fp0, x264cu_frame *fp1, int d0, int d1,
                                        float fps_factor, int referenced )
{
    if( !c || !fb || !fp0 || !fp1 || d0 < 1 || d1 < 0 || d0 + d1 > c->cfg.bframes + 1 ) FAIL( "x264cu_mbtree_propagate: bad argument" );
    if( !fb->evaluated[d0][d1] ) FAIL( "x264cu_mbtree_propagate: evaluation (%d,%d) has not been computed", d0, d1 );
    CU( cudaSetDevice( c->device ) );
    const LaParams &P = c->P;
    MbtreeJob J;
    J.intra = fb->costs;
    J.lowres_costs = fb->costs + (size_t)tri_index( c, d0, d1 ) * P.nmb;
    J.inv_qscale = fb->inv_qscale;
    J.prop_b = fb->prop;
    J.prop_ref[0] = fp0->prop; J.prop_ref[1] = fp1->prop;
    J.mvs[0] = fb->mvs + (size_t)field_index( c, 0, d0 ) * P.nmb * 2;
    J.mvs[1] = d1 ? fb->mvs + (size_t)field_index( c, 1, d1 ) * P.nmb * 2 : NULL;
    int dsf = ( ( d0 << 8 ) + ( ( d0 + d1 ) >> 1 ) ) / ( d0 + d1 );                  // slicetype.c:1054
    int w = P.weighted_bipred ? 64 - ( dsf >> 2 ) : 32;
    J.bipred_weight[0] = w; J.bipred_weight[1] = 64 - w;
    J.fps_factor = fps_factor;
    J.referenced = referenced;
    c->launches++;
    launch_mbtree_propagate( J, P, c->st );
    CU( cudaGetLastError() );
    return 0;
}

extern "C" int x264cu_mbtree_finish( x264cu_ctx *c, x264cu_frame *f, int fps_factor, float weightdelta, float strength,
                                     const float *qp_offset_aq, float *qp_offset )
{   // host arrays of mb_count floats; qp_offset must be initialised by the caller (blocks with intra_cost == 0 keep their value)
    if( !c || !f || !qp_offset ) FAIL( "x264cu_mbtree_finish: null argument" );
    if( !qp_offset_aq && !f->have_aq ) FAIL( "x264cu_mbtree_finish: no f_qp_offset_aq (pass it, or upload the frame with x264cu_frame_upload_aq)" );
    CU( cudaSetDevice( c->device ) );
    const size_t bytes = (size_t)c->P.nmb * 4;
    if( qp_offset_aq ) CU( cudaMemcpyAsync( f->qp_aq, qp_offset_aq, bytes, cudaMemcpyHostToDevice, c->st ) );
    CU( cudaMemcpyAsync( f->qp, qp_offset, bytes, cudaMemcpyHostToDevice, c->st ) );
    c->launches++;
    launch_mbtree_finish( f->costs, f->inv_qscale, f->prop, f->qp_aq, f->qp, c->d_log2_lut, fps_factor, weightdelta, strength, c->P.nmb, c->st );
    CU( cudaGetLastError() );
    CU( cudaMemcpyAsync( qp_offset, f->qp, bytes, cudaMemcpyDeviceToHost, c->st ) );
    CU( cudaStreamSynchronize( c->st ) );
    return 0;
}

extern "C" int x264cu_mbtree_get_propagate( x264cu_ctx *c, x264cu_frame *f, uint16_t *out )
{
    if( !c || !f || !out ) FAIL( "x264cu_mbtree_get_propagate: null argument" );
    CU( cudaSetDevice( c->device ) );
    std::vector<uint32_t> tmp( c->P.nmb );
    CU( cudaMemcpyAsync( tmp.data(), f->prop, (size_t)c->P.nmb * 4, cudaMemcpyDeviceToHost, c->st ) );
    CU( cudaStreamSynchronize( c->st ) );
    for( int i = 0; i < c->P.nmb; i++ ) out[i] = (uint16_t)( tmp[i] < 32767u ? tmp[i] : 32767u );
    return 0;
}

// ---- small helpers exported for callers that keep data resident ----
extern "C" void *x264cu_malloc( x264cu_ctx *c, size_t bytes )
{
    if( !c ) return NULL;
    void *p = NULL;
    cudaSetDevice( c->device );
    if( cudaMalloc( &p, bytes ) != cudaSuccess ) { snprintf( g_err, sizeof(g_err), "x264cu_malloc(%zu) failed", bytes ); cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void x264cu_free( x264cu_ctx *c, void *p ) { if( c && p ) { cudaSetDevice( c->device ); ctx_sync( c ); cudaFree( p ); } }
extern "C" void *x264cu_host_alloc( x264cu_ctx *c, size_t bytes )
{
    if( !c ) return NULL;
    void *p = NULL;
    cudaSetDevice( c->device );
    if( cudaHostAlloc( &p, bytes, cudaHostAllocDefault ) != cudaSuccess ) { snprintf( g_err, sizeof(g_err), "x264cu_host_alloc(%zu) failed", bytes ); cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void x264cu_host_free( x264cu_ctx *c, void *p ) { if( c && p ) { cudaSetDevice( c->device ); cudaFreeHost( p ); } }
extern "C" int x264cu_memcpy_h2d( x264cu_ctx *c, void *dst, const void *src, size_t bytes )
{
    if( !c ) FAIL( "null context" );
    CU( cudaSetDevice( c->device ) );
    CU( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyHostToDevice, c->st ) );
    CU( cudaStreamSynchronize( c->st ) );
    return 0;
}
extern "C" int x264cu_memcpy_d2d( x264cu_ctx *c, void *dst, const void *src, size_t bytes )
{
    if( !c ) FAIL( "null context" );
    CU( cudaSetDevice( c->device ) );
    CU( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToDevice, c->st ) );
    return 0;
}
extern "C" int x264cu_memcpy_d2h( x264cu_ctx *c, void *dst, const void *src, size_t bytes )
{
    if( !c ) FAIL( "null context" );
    CU( cudaSetDevice( c->device ) );
    CU( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToHost, c->st ) );
    CU( cudaStreamSynchronize( c->st ) );
    return 0;
}
extern "C" int x264cu_timer_begin( x264cu_ctx *c )
{
    if( !c ) FAIL( "null context" );
    CU( cudaSetDevice( c->device ) );
    CU( cudaEventRecord( c->tm_a, c->st ) );
    return 0;
}
extern "C" float x264cu_timer_end( x264cu_ctx *c )
{
    if( !c ) return -1.f;
    cudaSetDevice( c->device );
    float ms = -1.f;
    // everything enqueued so far, including result copies on the D2H stream
    if( cudaEventRecord( c->ev_res, c->st_out ) != cudaSuccess || cudaStreamWaitEvent( c->st, c->ev_res, 0 ) != cudaSuccess ) return -1.f;
    if( cudaEventRecord( c->tm_b, c->st ) != cudaSuccess ) return -1.f;
    if( cudaEventSynchronize( c->tm_b ) != cudaSuccess ) return -1.f;
    cudaEventElapsedTime( &ms, c->tm_a, c->tm_b );
    return ms;
}

// ------------------------------------------------------------------------------------------------
// batched primitives
// ------------------------------------------------------------------------------------------------
static const int g_bw[8] = { 16, 16, 8, 8, 8, 4, 4, 4 };
static const int g_bh[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

// RAII device temporaries for the host-pointer convenience path
struct DevTmp
{
    x264cu_ctx *c; std::vector<void *> ptrs; bool ok;
    DevTmp( x264cu_ctx *c_ ) : c( c_ ), ok( true ) {}
    void *up( const void *host, size_t bytes, size_t slack = 64 )
    {
        void *d = NULL;
        if( cudaMalloc( &d, bytes + slack ) != cudaSuccess ) { ok = false; return NULL; }
        ptrs.push_back( d );
        if( host && cudaMemcpyAsync( d, host, bytes, cudaMemcpyHostToDevice, c->st ) != cudaSuccess ) ok = false;
        return d;
    }
    int down( void *host, const void *dev, size_t bytes )
    {
        if( cudaMemcpyAsync( host, dev, bytes, cudaMemcpyDeviceToHost, c->st ) != cudaSuccess ) return -1;
        return cudaStreamSynchronize( c->st ) == cudaSuccess ? 0 : -1;
    }
    ~DevTmp() { cudaStreamSynchronize( c->st ); for( void *p : ptrs ) cudaFree( p ); }
};
#define PRIM_ENTER( c ) if( !(c) ) FAIL( "null context" ); CU( cudaSetDevice( (c)->device ) )
#define PRIM_CHECK( r, name ) do { if( (r) == -2 ) FAIL( name ": misaligned pointer or stride" ); if( (r) ) FAIL( name ": unsupported metric/size" ); \
                                   c->launches++; CU( cudaGetLastError() ); } while( 0 )

extern "C" int x264cu_pixel_cmp( x264cu_ctx *c, int metric, int size, const uint8_t *pix1, int64_t len1, int stride1,
                                 const uint8_t *pix2, int64_t len2, int stride2, int64_t n, const int64_t *off1,
                                 const int64_t *off2, int32_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( on_device )
    {
        int r = launch_pixel_cmp( metric, size, pix1, stride1, pix2, stride2, n, off1, off2, out, c->st );
        PRIM_CHECK( r, "x264cu_pixel_cmp" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *d1 = (uint8_t *)t.up( pix1, len1 ), *d2 = (uint8_t *)t.up( pix2, len2 );
    int64_t *o1 = (int64_t *)t.up( off1, n*8 ), *o2 = (int64_t *)t.up( off2, n*8 );
    int32_t *dout = (int32_t *)t.up( NULL, n*4 );
    if( !t.ok ) FAIL( "x264cu_pixel_cmp: device allocation/copy failed" );
    int r = launch_pixel_cmp( metric, size, d1, stride1, d2, stride2, n, o1, o2, dout, c->st );
    PRIM_CHECK( r, "x264cu_pixel_cmp" );
    if( t.down( out, dout, n*4 ) ) FAIL( "x264cu_pixel_cmp: copy back failed: %s", cudaGetErrorString( cudaGetLastError() ) );
    return 0;
}

extern "C" int x264cu_pixel_cmp_grid( x264cu_ctx *c, int metric, int size, const uint8_t *pix1, int stride1,
                                      const uint8_t *pix2, int stride2, int width, int height, int dx, int dy,
                                      int32_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( size < 0 || size > 7 || dx < 0 || dy < 0 ) FAIL( "x264cu_pixel_cmp_grid: bad argument" );
    int64_t nout = (int64_t)( width / g_bw[size] ) * ( height / g_bh[size] );
    if( on_device )
    {
        int r = launch_pixel_cmp_grid( metric, size, pix1, stride1, pix2 + (size_t)dy*stride2 + dx, stride2, width, height, out, c->st );
        PRIM_CHECK( r, "x264cu_pixel_cmp_grid" );
        return 0;
    }
    DevTmp t( c );
    // host planes are re-pitched to 256-byte aligned device rows
    int p1 = (int)align_up( width, 256 ), p2 = (int)align_up( width + dx, 256 );
    uint8_t *d1 = (uint8_t *)t.up( NULL, (size_t)p1 * height ), *d2 = (uint8_t *)t.up( NULL, (size_t)p2 * ( height + dy ) );
    int32_t *dout = (int32_t *)t.up( NULL, nout*4 );
    if( !t.ok ) FAIL( "x264cu_pixel_cmp_grid: device allocation failed" );
    CU( cudaMemcpy2DAsync( d1, p1, pix1, stride1, width, height, cudaMemcpyHostToDevice, c->st ) );
    CU( cudaMemcpy2DAsync( d2, p2, pix2, stride2, width + dx, height + dy, cudaMemcpyHostToDevice, c->st ) );
    int r = launch_pixel_cmp_grid( metric, size, d1, p1, d2 + (size_t)dy*p2 + dx, p2, width, height, dout, c->st );
    PRIM_CHECK( r, "x264cu_pixel_cmp_grid" );
    if( t.down( out, dout, nout*4 ) ) FAIL( "x264cu_pixel_cmp_grid: copy back failed" );
    return 0;
}

extern "C" int x264cu_sad_x( x264cu_ctx *c, int size, int nref, const uint8_t *fenc, const uint8_t *refs, int64_t reflen,
                             int stride, const int64_t *ref_off, int32_t *scores )
{
    PRIM_ENTER( c );
    if( nref < 1 || nref > 4 || size < 0 || size > 6 ) FAIL( "x264cu_sad_x: bad argument" );
    int64_t zero[4] = { 0, 0, 0, 0 };
    // fenc has FENC_STRIDE 16 (pixel.c:441-456)
    return x264cu_pixel_cmp( c, X264CU_SAD, size, fenc, 16*g_bh[size], 16, refs, reflen, stride, nref, zero, ref_off, scores, 0 );
}

extern "C" int x264cu_hadamard_ac( x264cu_ctx *c, int size, const uint8_t *pix, int64_t len, int stride, int64_t n,
                                   const int64_t *off, uint64_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( on_device )
    {
        int r = launch_hadamard_ac( size, pix, stride, n, off, out, c->st );
        PRIM_CHECK( r, "x264cu_hadamard_ac" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *d = (uint8_t *)t.up( pix, len );
    int64_t *o = (int64_t *)t.up( off, n*8 );
    uint64_t *dout = (uint64_t *)t.up( NULL, n*8 );
    if( !t.ok ) FAIL( "x264cu_hadamard_ac: device allocation/copy failed" );
    int r = launch_hadamard_ac( size, d, stride, n, o, dout, c->st );
    PRIM_CHECK( r, "x264cu_hadamard_ac" );
    if( t.down( out, dout, n*8 ) ) FAIL( "x264cu_hadamard_ac: copy back failed" );
    return 0;
}

extern "C" int x264cu_hadamard_ac_grid( x264cu_ctx *c, int size, const uint8_t *pix, int stride, int width, int height,
                                        uint64_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( size < 0 || size > 3 ) FAIL( "x264cu_hadamard_ac_grid: bad size" );
    int64_t nout = (int64_t)( width / g_bw[size] ) * ( height / g_bh[size] );
    if( on_device )
    {
        int r = launch_hadamard_ac_grid( size, pix, stride, width, height, out, c->st );
        PRIM_CHECK( r, "x264cu_hadamard_ac_grid" );
        return 0;
    }
    DevTmp t( c );
    int p = (int)align_up( width, 256 );
    uint8_t *d = (uint8_t *)t.up( NULL, (size_t)p * height );
    uint64_t *dout = (uint64_t *)t.up( NULL, nout*8 );
    if( !t.ok ) FAIL( "x264cu_hadamard_ac_grid: device allocation failed" );
    CU( cudaMemcpy2DAsync( d, p, pix, stride, width, height, cudaMemcpyHostToDevice, c->st ) );
    int r = launch_hadamard_ac_grid( size, d, p, width, height, dout, c->st );
    PRIM_CHECK( r, "x264cu_hadamard_ac_grid" );
    if( t.down( out, dout, nout*8 ) ) FAIL( "x264cu_hadamard_ac_grid: copy back failed" );
    return 0;
}

extern "C" int x264cu_sub_dct( x264cu_ctx *c, int bsize, const uint8_t *pix1, int64_t len1, int stride1,
                               const uint8_t *pix2, int64_t len2, int stride2, int64_t n, const int64_t *off1,
                               const int64_t *off2, int16_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( on_device )
    {
        int r = launch_sub_dct( bsize, pix1, stride1, pix2, stride2, n, off1, off2, out, c->st );
        PRIM_CHECK( r, "x264cu_sub_dct" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *d1 = (uint8_t *)t.up( pix1, len1 ), *d2 = (uint8_t *)t.up( pix2, len2 );
    int64_t *o1 = (int64_t *)t.up( off1, n*8 ), *o2 = (int64_t *)t.up( off2, n*8 );
    size_t ob = (size_t)n * bsize * bsize * 2;
    int16_t *dout = (int16_t *)t.up( NULL, ob );
    if( !t.ok ) FAIL( "x264cu_sub_dct: device allocation/copy failed" );
    int r = launch_sub_dct( bsize, d1, stride1, d2, stride2, n, o1, o2, dout, c->st );
    PRIM_CHECK( r, "x264cu_sub_dct" );
    if( t.down( out, dout, ob ) ) FAIL( "x264cu_sub_dct: copy back failed" );
    return 0;
}

extern "C" int x264cu_sub_dct_grid( x264cu_ctx *c, int bsize, const uint8_t *pix1, int stride1, const uint8_t *pix2,
                                    int stride2, int width, int height, int16_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( bsize != 4 && bsize != 8 && bsize != 16 ) FAIL( "x264cu_sub_dct_grid: bad block size" );
    size_t ob = (size_t)( width / bsize ) * ( height / bsize ) * bsize * bsize * 2;
    if( on_device )
    {
        int r = launch_sub_dct_grid( bsize, pix1, stride1, pix2, stride2, width, height, out, c->st );
        PRIM_CHECK( r, "x264cu_sub_dct_grid" );
        return 0;
    }
    DevTmp t( c );
    int p = (int)align_up( width, 256 );
    uint8_t *d1 = (uint8_t *)t.up( NULL, (size_t)p * height ), *d2 = (uint8_t *)t.up( NULL, (size_t)p * height );
    int16_t *dout = (int16_t *)t.up( NULL, ob );
    if( !t.ok ) FAIL( "x264cu_sub_dct_grid: device allocation failed" );
    CU( cudaMemcpy2DAsync( d1, p, pix1, stride1, width, height, cudaMemcpyHostToDevice, c->st ) );
    CU( cudaMemcpy2DAsync( d2, p, pix2, stride2, width, height, cudaMemcpyHostToDevice, c->st ) );
    int r = launch_sub_dct_grid( bsize, d1, p, d2, p, width, height, dout, c->st );
    PRIM_CHECK( r, "x264cu_sub_dct_grid" );
    if( t.down( out, dout, ob ) ) FAIL( "x264cu_sub_dct_grid: copy back failed" );
    return 0;
}

extern "C" int x264cu_quant( x264cu_ctx *c, int b8x8, int16_t *coefs, int64_t nblocks, const uint16_t *mf,
                             const uint16_t *bias, int32_t *nz, int on_device )
{
    PRIM_ENTER( c );
    int n = b8x8 ? 64 : 16;
    if( on_device )
    {
        int r = launch_quant( b8x8, coefs, nblocks, mf, bias, nz, c->st );
        PRIM_CHECK( r, "x264cu_quant" );
        return 0;
    }
    DevTmp t( c );
    int16_t *d = (int16_t *)t.up( coefs, (size_t)nblocks * n * 2 );
    uint16_t *dm = (uint16_t *)t.up( mf, n*2 ), *db = (uint16_t *)t.up( bias, n*2 );
    int32_t *dz = (int32_t *)t.up( NULL, nblocks*4 );
    if( !t.ok ) FAIL( "x264cu_quant: device allocation/copy failed" );
    int r = launch_quant( b8x8, d, nblocks, dm, db, dz, c->st );
    PRIM_CHECK( r, "x264cu_quant" );
    if( t.down( coefs, d, (size_t)nblocks * n * 2 ) || t.down( nz, dz, nblocks*4 ) ) FAIL( "x264cu_quant: copy back failed" );
    return 0;
}

extern "C" int x264cu_lowres_core( x264cu_ctx *c, const uint8_t *src, int src_stride, uint8_t *dst0, uint8_t *dsth,
                                   uint8_t *dstv, uint8_t *dstc, int dst_stride, int width, int height, int on_device )
{
    PRIM_ENTER( c );
    if( on_device )
    {
        int r = launch_lowres_core( src, src_stride, dst0, dsth, dstv, dstc, dst_stride, width, height, c->st );
        PRIM_CHECK( r, "x264cu_lowres_core" );
        return 0;
    }
    DevTmp t( c );
    size_t sb = (size_t)src_stride * ( 2*height + 1 ), db = (size_t)dst_stride * height;
    uint8_t *ds = (uint8_t *)t.up( src, sb );
    uint8_t *d[4]; for( int i = 0; i < 4; i++ ) d[i] = (uint8_t *)t.up( NULL, db );
    if( !t.ok ) FAIL( "x264cu_lowres_core: device allocation/copy failed" );
    for( int i = 0; i < 4; i++ ) CU( cudaMemsetAsync( d[i], 0, db, c->st ) );
    int r = launch_lowres_core( ds, src_stride, d[0], d[1], d[2], d[3], dst_stride, width, height, c->st );
    PRIM_CHECK( r, "x264cu_lowres_core" );
    uint8_t *h[4] = { dst0, dsth, dstv, dstc };
    for( int i = 0; i < 4; i++ ) if( t.down( h[i], d[i], db ) ) FAIL( "x264cu_lowres_core: copy back failed" );
    return 0;
}

// ---- checkasm surfaces of the fetch / prediction code (primitives_mc.cu) ----
extern "C" int x264cu_mc_luma( x264cu_ctx *c, const uint8_t *buf, int64_t len, const int64_t plane_off[4], int stride, int64_t n,
                               const int64_t *org, const int16_t *mv, int w, int h, const x264cu_weight *weight, uint8_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( !buf || !plane_off || ( n > 0 && ( !org || !mv || !out ) ) ) FAIL( "x264cu_mc_luma: null argument" );
    const int ws = weight && weight->enabled ? weight->scale : 0, wd = weight && weight->enabled ? weight->denom : -1, wo = weight && weight->enabled ? weight->offset : 0;
    if( on_device )
    {
        int r = launch_mc_luma( buf, plane_off, stride, n, org, mv, w, h, ws, wd, wo, out, c->st );
        PRIM_CHECK( r, "x264cu_mc_luma" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *d = (uint8_t *)t.up( buf, len );
    int64_t *o = (int64_t *)t.up( org, n*8 );
    int16_t *m = (int16_t *)t.up( mv, n*4 );
    uint8_t *dout = (uint8_t *)t.up( NULL, (size_t)n * w * h );
    if( !t.ok ) FAIL( "x264cu_mc_luma: device allocation/copy failed" );
    int r = launch_mc_luma( d, plane_off, stride, n, o, m, w, h, ws, wd, wo, dout, c->st );
    PRIM_CHECK( r, "x264cu_mc_luma" );
    if( n > 0 && t.down( out, dout, (size_t)n * w * h ) ) FAIL( "x264cu_mc_luma: copy back failed" );
    return 0;
}

static const int g_aw[12] = { 16, 16, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2 };      // PIXEL_16x16 .. PIXEL_2x2 (pixel.h:52-68)
static const int g_ah[12] = { 16, 8, 16, 8, 4, 8, 4, 16, 2, 8, 4, 2 };

extern "C" int x264cu_pixel_avg( x264cu_ctx *c, int size, const uint8_t *pix1, int64_t len1, int stride1, const uint8_t *pix2, int64_t len2, int stride2,
                                 int64_t n, const int64_t *off1, const int64_t *off2, int weight, uint8_t *out, int on_device )
{
    PRIM_ENTER( c );
    if( size < 0 || size > 11 ) FAIL( "x264cu_pixel_avg: bad size" );
    const int w = g_aw[size], h = g_ah[size];
    if( on_device )
    {
        int r = launch_pixel_avg( pix1, stride1, pix2, stride2, n, off1, off2, w, h, weight, out, c->st );
        PRIM_CHECK( r, "x264cu_pixel_avg" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *d1 = (uint8_t *)t.up( pix1, len1 ), *d2 = (uint8_t *)t.up( pix2, len2 );
    int64_t *o1 = (int64_t *)t.up( off1, n*8 ), *o2 = (int64_t *)t.up( off2, n*8 );
    uint8_t *dout = (uint8_t *)t.up( NULL, (size_t)n * w * h );
    if( !t.ok ) FAIL( "x264cu_pixel_avg: device allocation/copy failed" );
    int r = launch_pixel_avg( d1, stride1, d2, stride2, n, o1, o2, w, h, weight, dout, c->st );
    PRIM_CHECK( r, "x264cu_pixel_avg" );
    if( n > 0 && t.down( out, dout, (size_t)n * w * h ) ) FAIL( "x264cu_pixel_avg: copy back failed" );
    return 0;
}

extern "C" int x264cu_mc_weight( x264cu_ctx *c, const uint8_t *src, int src_stride, uint8_t *dst, int dst_stride, int width, int height,
                                 const x264cu_weight *weight, int on_device )
{
    PRIM_ENTER( c );
    if( !src || !dst || !weight || weight->denom < 0 ) FAIL( "x264cu_mc_weight: bad argument" );
    if( on_device )
    {
        int r = launch_mc_weight( src, src_stride, dst, dst_stride, width, height, weight->scale, weight->denom, weight->offset, c->st );
        PRIM_CHECK( r, "x264cu_mc_weight" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *ds = (uint8_t *)t.up( src, (size_t)src_stride * height ), *dd = (uint8_t *)t.up( NULL, (size_t)width * height );
    if( !t.ok ) FAIL( "x264cu_mc_weight: device allocation/copy failed" );
    int r = launch_mc_weight( ds, src_stride, dd, width, width, height, weight->scale, weight->denom, weight->offset, c->st );
    PRIM_CHECK( r, "x264cu_mc_weight" );
    CU( cudaMemcpy2DAsync( dst, dst_stride, dd, width, width, height, cudaMemcpyDeviceToHost, c->st ) );
    CU( cudaStreamSynchronize( c->st ) );
    return 0;
}

extern "C" int x264cu_intra_x3_8x8c( x264cu_ctx *c, int b_satd, const uint8_t *fenc, int64_t len1, int stride1, const uint8_t *fdec, int64_t len2, int stride2,
                                     int64_t n, const int64_t *off1, const int64_t *off2, int32_t *res, int on_device )
{
    PRIM_ENTER( c );
    if( on_device )
    {
        int r = launch_intra_x3_8x8c( b_satd, fenc, stride1, fdec, stride2, n, off1, off2, res, c->st );
        PRIM_CHECK( r, "x264cu_intra_x3_8x8c" );
        return 0;
    }
    DevTmp t( c );
    uint8_t *d1 = (uint8_t *)t.up( fenc, len1 ), *d2 = (uint8_t *)t.up( fdec, len2 );
    int64_t *o1 = (int64_t *)t.up( off1, n*8 ), *o2 = (int64_t *)t.up( off2, n*8 );
    int32_t *dres = (int32_t *)t.up( NULL, (size_t)n * 12 );
    if( !t.ok ) FAIL( "x264cu_intra_x3_8x8c: device allocation/copy failed" );
    int r = launch_intra_x3_8x8c( b_satd, d1, stride1, d2, stride2, n, o1, o2, dres, c->st );
    PRIM_CHECK( r, "x264cu_intra_x3_8x8c" );
    if( n > 0 && t.down( res, dres, (size_t)n * 12 ) ) FAIL( "x264cu_intra_x3_8x8c: copy back failed" );
    return 0;
}

extern "C" int x264cu_quant_4x4x4( x264cu_ctx *c, int16_t *coefs, int64_t ngroups, const uint16_t *mf, const uint16_t *bias, int32_t *nz_mask, int on_device )
{
    PRIM_ENTER( c );
    if( on_device )
    {
        int r = launch_quant_4x4x4( coefs, ngroups, mf, bias, nz_mask, c->st );
        PRIM_CHECK( r, "x264cu_quant_4x4x4" );
        return 0;
    }
    DevTmp t( c );
    int16_t *d = (int16_t *)t.up( coefs, (size_t)ngroups * 64 * 2 );
    uint16_t *dm = (uint16_t *)t.up( mf, 32 ), *db = (uint16_t *)t.up( bias, 32 );
    int32_t *dz = (int32_t *)t.up( NULL, ngroups*4 );
    if( !t.ok ) FAIL( "x264cu_quant_4x4x4: device allocation/copy failed" );
    int r = launch_quant_4x4x4( d, ngroups, dm, db, dz, c->st );
    PRIM_CHECK( r, "x264cu_quant_4x4x4" );
    if( ngroups > 0 && ( t.down( coefs, d, (size_t)ngroups * 64 * 2 ) || t.down( nz_mask, dz, ngroups*4 ) ) ) FAIL( "x264cu_quant_4x4x4: copy back failed" );
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Peer exchange for one encode stream sharded over several GPUs (the MV field a B evaluation
// needs from the rank that searched it).  Every rank owns a mailbox in its own HBM, exported with CUDA IPC; the
// owner of a field stores it straight into every peer's mailbox over NVLink from ONE kernel (no copy-engine launches,
// no collective, no host synchronisation) and then raises a flag in the peer's mailbox header; the consumer's import
// kernel waits for the flag on the device and copies mailbox -> frame fields.  Mailboxes are PEER_DEPTH deep: a rank's
// push of step m+4 follows (in its stream) its import of step m+2, which waited for this rank's push of step m+2,
// which follows this rank's import of step m -- so a slot is never overwritten before it was read.
// ------------------------------------------------------------------------------------------------
enum { PEER_DEPTH = 4, PEER_MAX_ITEMS = 64, PEER_MAX_WORLD = 16, PEER_HEADER = 256 };

struct PeerItems
{
    const uint32_t *src[PEER_MAX_ITEMS];    // push: local field; import: unused
    uint32_t *dst[PEER_MAX_ITEMS];          // import: local field
    short from[PEER_MAX_ITEMS], slot[PEER_MAX_ITEMS];
    int n, words;
};
struct PeerTargets { uint8_t *box[PEER_MAX_WORLD]; int world, rank; };

struct x264cu_peer
{
    x264cu_ctx *c;
    int rank, world, items_per_rank;
    size_t slot_stride, bytes;
    uint8_t *box;                           // local mailbox: header (flags[world], error) + [PEER_DEPTH][world][items_per_rank][slot_stride]
    uint8_t *remote[PEER_MAX_WORLD];        // peers' mailboxes as mapped here (remote[rank] = box)
    void *remote_base[PEER_MAX_WORLD];
    unsigned *d_done;                       // blocks of all push launches that finished their stores
    unsigned long long done_expected;
    uint32_t push_epoch, pull_epoch;
    int *h_err;
};

static __device__ __forceinline__ void st_release_sys( uint32_t *p, uint32_t v ) { asm volatile( "st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory" ); }
static __device__ __forceinline__ uint32_t ld_acquire_sys( const uint32_t *p )
{ uint32_t v; asm volatile( "ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory" ); return v; }

static __device__ __forceinline__ size_t peer_slot_offset( int depth_slot, int world, int items_per_rank, size_t slot_stride, int from, int slot )
{ return PEER_HEADER + ( ( (size_t)depth_slot * world + from ) * items_per_rank + slot ) * slot_stride; }

// grid (chunks, max(items,1)): every block stores its chunk of one field into all peers' mailboxes; the last block to
// finish (device-wide counter) publishes the step to every peer
static __global__ void __launch_bounds__( 256 ) peer_push_kernel( const __grid_constant__ PeerItems it, const __grid_constant__ PeerTargets tg,
                                                                  int items_per_rank, size_t slot_stride, uint32_t epoch,
                                                                  unsigned *done, unsigned done_expected )
{
    const int item = blockIdx.y;
    if( item < it.n )
    {
        const uint32_t *src = it.src[item];
        const size_t off = peer_slot_offset( epoch % PEER_DEPTH, tg.world, items_per_rank, slot_stride, tg.rank, it.slot[item] );
        for( int w = blockIdx.x * blockDim.x + threadIdx.x; w < it.words; w += gridDim.x * blockDim.x )
        {
            const uint32_t v = src[w];
            for( int p = 0; p < tg.world; p++ )
                if( p != tg.rank ) ( (uint32_t *)( tg.box[p] + off ) )[w] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if( threadIdx.x == 0 && atomicAdd( done, 1u ) + 1 == done_expected )
    {
        __threadfence_system();
        for( int p = 0; p < tg.world; p++ )
            if( p != tg.rank ) st_release_sys( (uint32_t *)tg.box[p] + tg.rank, epoch + 1 );
    }
}

// grid (chunks, max(items,1)): wait until the source rank has published the step, then mailbox -> frame field.
// Block (0,0) also waits for every other peer, so that the mailbox re-use argument above holds for ranks that own nothing.
static __global__ void __launch_bounds__( 256 ) peer_import_kernel( const __grid_constant__ PeerItems it, uint8_t *box, int world, int rank,
                                                                    int items_per_rank, size_t slot_stride, uint32_t epoch )
{
    const int item = blockIdx.y;
    uint32_t *flags = (uint32_t *)box;
    __shared__ int ok;
    if( threadIdx.x == 0 )
    {
        ok = 1;
        const bool all = blockIdx.x == 0 && blockIdx.y == 0;
        for( int p = 0; p < world; p++ )
        {
            if( p == rank || ( !all && ( item >= it.n || p != it.from[item] ) ) ) continue;
            unsigned ns = 64;
            long long t0 = clock64();
            while( (int32_t)( ld_acquire_sys( flags + p ) - ( epoch + 1 ) ) < 0 )
            {
                __nanosleep( ns ); if( ns < 2048 ) ns *= 2;
                if( clock64() - t0 > 20000000000ll ) { flags[PEER_MAX_WORLD] = 1; ok = 0; break; }     // ~10 s: a peer died
            }
        }
    }
    __syncthreads();
    if( !ok || item >= it.n ) return;
    const uint32_t *src = (const uint32_t *)( box + peer_slot_offset( epoch % PEER_DEPTH, world, items_per_rank, slot_stride, it.from[item], it.slot[item] ) );
    uint32_t *dst = it.dst[item];
    for( int w = blockIdx.x * blockDim.x + threadIdx.x; w < it.words; w += gridDim.x * blockDim.x )
        dst[w] = __ldcg( src + w );
}

extern "C" x264cu_peer *x264cu_peer_create( x264cu_ctx *c, int rank, int world, int items_per_rank, uint8_t handle_out[64], uint64_t *offset_out )
{
    if( !c || world < 2 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || items_per_rank < 1 || items_per_rank > PEER_MAX_ITEMS || !handle_out || !offset_out )
    { snprintf( g_err, sizeof(g_err), "x264cu_peer_create: bad argument" ); return NULL; }
    CUP( cudaSetDevice( c->device ) );
    x264cu_peer *p = new x264cu_peer();
    memset( p, 0, sizeof(*p) );
    p->c = c; p->rank = rank; p->world = world; p->items_per_rank = items_per_rank;
    p->slot_stride = ( (size_t)c->P.nmb * 4 + 255 ) & ~(size_t)255;
    p->bytes = PEER_HEADER + (size_t)PEER_DEPTH * world * items_per_rank * p->slot_stride;
    if( p->bytes < ( 4u << 20 ) ) p->bytes = 4u << 20;      // its own allocation block, not a sub-allocation
    cudaIpcMemHandle_t h;
    if( cudaMalloc( &p->box, p->bytes ) != cudaSuccess || cudaMemset( p->box, 0, PEER_HEADER ) != cudaSuccess ||
        cudaMalloc( &p->d_done, 4 ) != cudaSuccess || cudaMemset( p->d_done, 0, 4 ) != cudaSuccess ||
        cudaIpcGetMemHandle( &h, p->box ) != cudaSuccess )
    {
        snprintf( g_err, sizeof(g_err), "x264cu_peer_create: %s", cudaGetErrorString( cudaGetLastError() ) );
        cudaFree( p->box ); cudaFree( p->d_done ); delete p; return NULL;
    }
    // offset of the mailbox inside the allocation the handle names (0 unless the allocator sub-allocated)
    uint64_t off = 0;
    {
        typedef int (*range_fn)( unsigned long long *, size_t *, unsigned long long );
        void *fn = NULL; cudaDriverEntryPointQueryResult qr;
        if( cudaGetDriverEntryPoint( "cuMemGetAddressRange", &fn, cudaEnableDefault, &qr ) == cudaSuccess && fn )
        {
            unsigned long long base = 0; size_t sz = 0;
            if( ( (range_fn)fn )( &base, &sz, (unsigned long long)(uintptr_t)p->box ) == 0 ) off = (uint64_t)(uintptr_t)p->box - base;
        }
    }
    static_assert( sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes" );
    memcpy( handle_out, &h, 64 ); *offset_out = off;
    p->remote[rank] = p->box;
    return p;
}

extern "C" int x264cu_peer_connect( x264cu_peer *p, const uint8_t *handles, const uint64_t *offsets )
{
    if( !p || !handles || !offsets ) FAIL( "x264cu_peer_connect: null argument" );
    CU( cudaSetDevice( p->c->device ) );
    for( int r = 0; r < p->world; r++ )
    {
        if( r == p->rank ) continue;
        cudaIpcMemHandle_t h; memcpy( &h, handles + (size_t)r * 64, 64 );
        CU( cudaIpcOpenMemHandle( &p->remote_base[r], h, cudaIpcMemLazyEnablePeerAccess ) );
        p->remote[r] = (uint8_t *)p->remote_base[r] + offsets[r];
    }
    return 0;
}

// ranks that live in ONE process (one encoder driving several GPUs from its lookahead thread, or several contexts on one
// device): no IPC handle is needed -- or allowed: a handle cannot be opened by the process that exported it -- the peers'
// mailboxes are mapped by enabling peer access between the devices
extern "C" int x264cu_peer_connect_local( x264cu_peer *p, x264cu_peer *const *all, int world )
{
    if( !p || !all || world != p->world ) FAIL( "x264cu_peer_connect_local: bad argument" );
    CU( cudaSetDevice( p->c->device ) );
    for( int r = 0; r < world; r++ )
    {
        if( r == p->rank ) continue;
        if( !all[r] || all[r]->rank != r || all[r]->world != world || all[r]->items_per_rank != p->items_per_rank || all[r]->slot_stride != p->slot_stride )
            FAIL( "x264cu_peer_connect_local: peer %d does not match", r );
        const int dev = all[r]->c->device;
        if( dev != p->c->device )
        {
            int can = 0;
            CU( cudaDeviceCanAccessPeer( &can, p->c->device, dev ) );
            if( !can ) FAIL( "x264cu_peer_connect_local: device %d cannot access device %d", p->c->device, dev );
            cudaError_t e = cudaDeviceEnablePeerAccess( dev, 0 );
            if( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled ) CU( e );
            cudaGetLastError();
        }
        p->remote[r] = all[r]->box;
        p->remote_base[r] = NULL;       // nothing to close
    }
    return 0;
}

static int peer_fill( x264cu_peer *p, PeerItems &it, const void *const *ptrs, const int *from, const int *slot, int n, bool push )
{
    if( n < 0 || n > PEER_MAX_ITEMS ) return -1;
    memset( &it, 0, sizeof(it) );
    it.n = n; it.words = p->c->P.nmb;
    for( int i = 0; i < n; i++ )
    {
        if( slot[i] < 0 || slot[i] >= p->items_per_rank || ( !push && ( from[i] < 0 || from[i] >= p->world || from[i] == p->rank ) ) ) return -1;
        if( push ) it.src[i] = (const uint32_t *)ptrs[i]; else it.dst[i] = (uint32_t *)ptrs[i];
        it.from[i] = (short)( push ? p->rank : from[i] ); it.slot[i] = (short)slot[i];
    }
    return 0;
}

// enqueue on the context's stream: fields src[i] (nmb 32-bit words each) -> slot[i] of this rank's row in every peer's mailbox
extern "C" int x264cu_peer_push( x264cu_peer *p, const void *const *src, const int *slot, int n )
{
    if( !p ) FAIL( "x264cu_peer_push: null argument" );
    x264cu_ctx *c = p->c;
    CU( cudaSetDevice( c->device ) );
    PeerItems it;
    if( peer_fill( p, it, src, NULL, slot, n, true ) ) FAIL( "x264cu_peer_push: bad item list" );
    PeerTargets tg; memset( &tg, 0, sizeof(tg) );
    tg.world = p->world; tg.rank = p->rank;
    for( int r = 0; r < p->world; r++ ) { tg.box[r] = p->remote[r]; if( !tg.box[r] ) FAIL( "x264cu_peer_push: not connected" ); }
    int chunks = ( it.words + 2047 ) / 2048; if( chunks < 1 ) chunks = 1;
    dim3 grid( chunks, n > 0 ? n : 1 );
    p->done_expected += (unsigned long long)grid.x * grid.y;
    peer_push_kernel<<<grid, 256, 0, c->st>>>( it, tg, p->items_per_rank, p->slot_stride, p->push_epoch, p->d_done, (unsigned)p->done_expected );
    CU( cudaGetLastError() );
    c->launches++;
    p->push_epoch++;
    return 0;
}

// enqueue on the context's stream: wait (on the device) for the peers' push of the same step, then slot[i] of rank from[i] -> dst[i]
extern "C" int x264cu_peer_import( x264cu_peer *p, void *const *dst, const int *from, const int *slot, int n )
{
    if( !p ) FAIL( "x264cu_peer_import: null argument" );
    x264cu_ctx *c = p->c;
    CU( cudaSetDevice( c->device ) );
    PeerItems it;
    if( peer_fill( p, it, (const void *const *)dst, from, slot, n, false ) ) FAIL( "x264cu_peer_import: bad item list" );
    int chunks = ( it.words + 2047 ) / 2048; if( chunks < 1 ) chunks = 1;
    dim3 grid( chunks, n > 0 ? n : 1 );
    peer_import_kernel<<<grid, 256, 0, c->st>>>( it, p->box, p->world, p->rank, p->items_per_rank, p->slot_stride, p->pull_epoch );
    CU( cudaGetLastError() );
    c->launches++;
    p->pull_epoch++;
    return 0;
}

// after a flush: 1 when an import gave up waiting for a peer
extern "C" int x264cu_peer_error( x264cu_peer *p )
{
    if( !p ) return -1;
    uint32_t e = 0;
    if( cudaSetDevice( p->c->device ) != cudaSuccess ||
        cudaMemcpy( &e, (uint32_t *)p->box + PEER_MAX_WORLD, 4, cudaMemcpyDeviceToHost ) != cudaSuccess ) return -1;
    return (int)e;
}

extern "C" void x264cu_peer_destroy( x264cu_peer *p )
{
    if( !p ) return;
    cudaSetDevice( p->c->device );
    cudaStreamSynchronize( p->c->st );
    for( int r = 0; r < p->world; r++ ) if( r != p->rank && p->remote_base[r] ) cudaIpcCloseMemHandle( p->remote_base[r] );
    cudaFree( p->box ); cudaFree( p->d_done );
    delete p;
}
