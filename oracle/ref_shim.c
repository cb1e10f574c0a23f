/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A flat C surface over the *unmodified* reference sources under /root/reference, compiled by
 * oracle/Makefile into oracle/_ref/libx264ref.so.  This translation unit is compiled IN PLACE OF
 * encoder/analyse.c: it textually includes that file (which itself includes rdo.c and
 * slicetype.c, analyse.c:33,3894) so that the file-static lookahead functions
 * (slicetype_mb_cost slicetype.c:514, slicetype_frame_cost :836, lowres_context_init :45) are
 * reachable without touching the reference tree.  No reference source is copied into this repo.
 *
 * Two surfaces:
 *   ref_*      primitive tables exactly as x264_*_init(cpu=0) fills them (pixel.c:809, mc.c:613,
 *              dct.c:477, quant.c:414, predict.c) -- the checkasm "C reference" side.
 *   ref_la_*   the lookahead: real x264_t from x264_encoder_open, real x264_frame_t with
 *              x264_frame_init_lowres, then either the reference's own raster-serial
 *              slicetype_frame_cost, or the dependency-free ("Jacobi") schedule driving the
 *              UNMODIFIED slicetype_mb_cost block by block with neighbour MVs restored from the
 *              previous pass's snapshot.  The latter pins oracle/x264_oracle.c.
 */
#include "encoder/analyse.c"

#include <time.h>

/* encoder.c defines these templated entry points; api.c declares them privately (api.c:34-40) */
x264_t *x264_encoder_open( x264_param_t *, void * );
void    x264_encoder_close( x264_t * );

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* primitive tables                                                                            */
/* ------------------------------------------------------------------------------------------ */
static x264_pixel_function_t  g_pixf;
static x264_mc_functions_t    g_mc;
static x264_dct_function_t    g_dct;
static x264_quant_function_t  g_qf;
static x264_predict_t         g_pred8x8c[7];
static x264_predict8x8_t      g_pred8x8[12];
static x264_predict_8x8_filter_t g_pred8x8_filter;
static x264_t                *g_h;      /* small encoder instance: source of CQM tables */
static int                    g_ready;

static void quiet_log( void *p, int level, const char *fmt, va_list ap ) { (void)p; (void)level; (void)fmt; (void)ap; }

static x264_t *open_encoder( int width, int height, int bframes, int subme, int mbtree, int aq_mode,
                             int weightp, int b_adapt, int la_threads, int rc_lookahead, int vbv )
{
    x264_param_t param;
    x264_param_default( &param );           /* == --preset medium (base.c:553-556) */
    param.i_width = width;
    param.i_height = height;
    param.i_csp = X264_CSP_I420;
    param.i_bitdepth = 8;
    /* lookahead slices need i_threads > 1 (encoder.c:580-584); sliced threads keep i_thread_frames == 1 */
    param.i_threads = la_threads > 1 ? la_threads : 1;
    param.i_lookahead_threads = la_threads;
    param.b_sliced_threads = la_threads > 1;
    param.i_sync_lookahead = 0;
    param.i_bframe = bframes;
    param.i_bframe_adaptive = b_adapt;
    param.analyse.i_subpel_refine = subme;
    param.analyse.i_weighted_pred = weightp;
    param.rc.b_mb_tree = mbtree;
    param.rc.i_aq_mode = aq_mode;
    param.rc.i_lookahead = rc_lookahead;
    param.i_fps_num = 30; param.i_fps_den = 1;
    param.b_vfr_input = 0;
    param.b_opencl = 0;
    param.pf_log = quiet_log;
    param.cpu = 0;
    if( vbv )
    {
        param.rc.i_vbv_buffer_size = vbv;
        param.rc.i_vbv_max_bitrate = vbv;
    }
    return x264_encoder_open( &param, NULL );   /* templated: x264_8_encoder_open */
}

static void ensure_tables( void )
{
    if( g_ready ) return;
    x264_pixel_init( 0, &g_pixf );
    x264_mc_init( 0, &g_mc, 1 );
    x264_dct_init( 0, &g_dct );
    x264_predict_8x8c_init( 0, g_pred8x8c );
    x264_predict_8x8_init( 0, g_pred8x8, &g_pred8x8_filter );
    g_h = open_encoder( 64, 64, 0, 7, 0, 1, 0, 0, 1, 0, 0 );
    x264_quant_init( g_h, 0, &g_qf );
    g_ready = 1;
}

/* kind: 0 sad, 1 satd, 2 sa8d, 3 ssd.  size: PIXEL_16x16..PIXEL_4x4 (pixel.h:41-52 order:
 * 16x16,16x8,8x16,8x8,8x4,4x8,4x4,4x16). */
static x264_pixel_cmp_t pick_cmp( int kind, int size )
{
    ensure_tables();
    switch( kind )
    {
        case 0: return g_pixf.sad[size];
        case 1: return g_pixf.satd[size];
        case 2: return g_pixf.sa8d[size];
        case 3: return g_pixf.ssd[size];
    }
    return NULL;
}

EXPORT int ref_pixel_cmp( int kind, int size, uint8_t *p1, intptr_t s1, uint8_t *p2, intptr_t s2 )
{
    x264_pixel_cmp_t f = pick_cmp( kind, size );
    return f ? f( p1, s1, p2, s2 ) : -1;
}

EXPORT int ref_pixel_cmp_batch( int kind, int size, uint8_t *base1, intptr_t s1, uint8_t *base2, intptr_t s2,
                                int64_t n, const int64_t *off1, const int64_t *off2, int32_t *out )
{
    x264_pixel_cmp_t f = pick_cmp( kind, size );
    if( !f ) return -1;
    for( int64_t i = 0; i < n; i++ )
        out[i] = f( base1 + off1[i], s1, base2 + off2[i], s2 );
    return 0;
}

EXPORT uint64_t ref_hadamard_ac( int size, uint8_t *pix, intptr_t stride )
{
    ensure_tables();
    return g_pixf.hadamard_ac[size] ? g_pixf.hadamard_ac[size]( pix, stride ) : ~0ULL;
}

EXPORT int ref_hadamard_ac_batch( int size, uint8_t *base, intptr_t stride, int64_t n, const int64_t *off, uint64_t *out )
{
    ensure_tables();
    if( !g_pixf.hadamard_ac[size] ) return -1;
    for( int64_t i = 0; i < n; i++ )
        out[i] = g_pixf.hadamard_ac[size]( base + off[i], stride );
    return 0;
}

/* fenc has FENC_STRIDE (16) as in pixel.c:441-464 */
EXPORT void ref_sad_x3( int size, uint8_t *fenc, uint8_t *p0, uint8_t *p1, uint8_t *p2, intptr_t stride, int *scores )
{
    ensure_tables();
    g_pixf.sad_x3[size]( fenc, p0, p1, p2, stride, scores );
}
EXPORT void ref_sad_x4( int size, uint8_t *fenc, uint8_t *p0, uint8_t *p1, uint8_t *p2, uint8_t *p3, intptr_t stride, int *scores )
{
    ensure_tables();
    g_pixf.sad_x4[size]( fenc, p0, p1, p2, p3, stride, scores );
}

/* fdec: pointer to the 8x8 block inside a FDEC_STRIDE(32) scratch with neighbours in place */
EXPORT void ref_intra_x3_8x8c( int b_satd, uint8_t *fenc, uint8_t *fdec, int *res )
{
    ensure_tables();
    if( b_satd ) g_pixf.intra_satd_x3_8x8c( fenc, fdec, res );
    else         g_pixf.intra_sad_x3_8x8c( fenc, fdec, res );
}
EXPORT void ref_predict_8x8c( int mode, uint8_t *src ) { ensure_tables(); g_pred8x8c[mode]( src ); }
EXPORT void ref_predict_8x8_filter( uint8_t *src, uint8_t *edge, int neighbors, int filters )
{ ensure_tables(); g_pred8x8_filter( src, edge, neighbors, filters ); }
EXPORT void ref_predict_8x8( int mode, uint8_t *src, uint8_t *edge ) { ensure_tables(); g_pred8x8[mode]( src, edge ); }

/* DCT: pix1 stride FENC_STRIDE(16), pix2 stride FDEC_STRIDE(32) (dct.h:30-32) */
EXPORT void ref_sub4x4_dct( int16_t *dct, uint8_t *pix1, uint8_t *pix2 ) { ensure_tables(); g_dct.sub4x4_dct( dct, pix1, pix2 ); }
EXPORT void ref_sub8x8_dct( int16_t *dct, uint8_t *pix1, uint8_t *pix2 ) { ensure_tables(); g_dct.sub8x8_dct( (dctcoef(*)[16])dct, pix1, pix2 ); }
EXPORT void ref_sub16x16_dct( int16_t *dct, uint8_t *pix1, uint8_t *pix2 ) { ensure_tables(); g_dct.sub16x16_dct( (dctcoef(*)[16])dct, pix1, pix2 ); }

/* batched over two planes with arbitrary strides: blocks are staged into FENC/FDEC-strided
 * scratch, then the reference function is called. bsize = 4, 8 or 16. */
EXPORT int ref_sub_dct_batch( int bsize, uint8_t *base1, intptr_t s1, uint8_t *base2, intptr_t s2,
                              int64_t n, const int64_t *off1, const int64_t *off2, int16_t *out )
{
    ensure_tables();
    ALIGNED_ARRAY_64( uint8_t, fenc,[16*FENC_STRIDE] );
    ALIGNED_ARRAY_64( uint8_t, fdec,[16*FDEC_STRIDE] );
    ALIGNED_ARRAY_64( dctcoef, coef,[16],[16] );
    int ncoef = bsize*bsize;
    for( int64_t i = 0; i < n; i++ )
    {
        for( int y = 0; y < bsize; y++ )
        {
            memcpy( fenc + y*FENC_STRIDE, base1 + off1[i] + y*s1, bsize );
            memcpy( fdec + y*FDEC_STRIDE, base2 + off2[i] + y*s2, bsize );
        }
        if( bsize == 4 )       g_dct.sub4x4_dct( coef[0], fenc, fdec );
        else if( bsize == 8 )  g_dct.sub8x8_dct( coef, fenc, fdec );
        else if( bsize == 16 ) g_dct.sub16x16_dct( coef, fenc, fdec );
        else return -1;
        memcpy( out + i*ncoef, coef, ncoef*sizeof(dctcoef) );
    }
    return 0;
}

EXPORT int ref_quant_4x4( int16_t *dct, uint16_t *mf, uint16_t *bias ) { ensure_tables(); return g_qf.quant_4x4( dct, mf, bias ); }
EXPORT int ref_quant_8x8( int16_t *dct, uint16_t *mf, uint16_t *bias ) { ensure_tables(); return g_qf.quant_8x8( dct, mf, bias ); }
EXPORT int ref_quant_4x4x4( int16_t *dct, uint16_t *mf, uint16_t *bias ) { ensure_tables(); return g_qf.quant_4x4x4( (dctcoef(*)[16])dct, mf, bias ); }
EXPORT int ref_quant_batch( int b8x8, int16_t *coefs, int64_t nblocks, uint16_t *mf, uint16_t *bias, int32_t *nz )
{
    ensure_tables();
    int n = b8x8 ? 64 : 16;
    for( int64_t i = 0; i < nblocks; i++ )
        nz[i] = b8x8 ? g_qf.quant_8x8( coefs + i*n, mf, bias ) : g_qf.quant_4x4( coefs + i*n, mf, bias );
    return 0;
}

/* flat-CQM quant tables produced by x264_cqm_init (set.c:73). cat: 0 intra-Y, 1 inter-Y. */
EXPORT int ref_cqm_tables( int b8x8, int cat, int qp, uint16_t *mf, uint16_t *bias )
{
    ensure_tables();
    if( qp < 0 || qp > QP_MAX_SPEC ) return -1;
    if( b8x8 )
    {
        int c = cat ? CQM_8PY : CQM_8IY;
        memcpy( mf, g_h->quant8_mf[c][qp], 64*sizeof(uint16_t) );
        memcpy( bias, g_h->quant8_bias[c][qp], 64*sizeof(uint16_t) );
    }
    else
    {
        int c = cat ? CQM_4PY : CQM_4IY;
        memcpy( mf, g_h->quant4_mf[c][qp], 16*sizeof(uint16_t) );
        memcpy( bias, g_h->quant4_bias[c][qp], 16*sizeof(uint16_t) );
    }
    return 0;
}

/* the six CQM configurations of check_quant (tools/checkasm.c:2098-2139: flat, JVT, custom 4x4, custom 8x8, random, all-ones)
 * installed on the table-source encoder; ref_cqm_tables then returns their tables.  The custom lists are generated here
 * from `seed` (any list exercises the custom path).  i_cqm < 0 restores the flat default. */
static uint8_t g_cqm_buf[3][64];
EXPORT int ref_cqm_select( int i_cqm, unsigned seed )
{
    ensure_tables();
    x264_t *h = g_h;
    x264_cqm_delete( h );
    srand( seed );
    for( int i = 0; i < 64; i++ )
    {
        g_cqm_buf[0][i] = 16 + 8 * ( ( ( i & 3 ) == 2 ) + ( ( ( i >> 2 ) & 3 ) == 2 ) ) + rand() % 4;   /* a mild custom 4x4 list */
        g_cqm_buf[1][i] = 3 + ( ( i & 3 ) == 2 ) + ( ( ( i >> 3 ) & 3 ) == 2 ) + rand() % 2;            /* a steep custom 8x8 list */
        g_cqm_buf[2][i] = i_cqm == 5 ? 1 : 10 + rand() % 246;                                          /* random / all-ones */
    }
    int preset = X264_CQM_CUSTOM;
    for( int i = 0; i < 8; i++ )
        switch( i_cqm )
        {
            case 1:  h->sps->scaling_list[i] = x264_cqm_jvt[i]; preset = X264_CQM_JVT; break;
            case 2:  h->sps->scaling_list[i] = i < 4 ? g_cqm_buf[0] : x264_cqm_flat16; break;
            case 3:  h->sps->scaling_list[i] = i < 4 ? x264_cqm_flat16 : g_cqm_buf[1]; break;
            case 4: case 5: h->sps->scaling_list[i] = g_cqm_buf[2]; break;
            default: h->sps->scaling_list[i] = x264_cqm_flat16; preset = X264_CQM_FLAT; break;
        }
    h->param.i_cqm_preset = h->sps->i_cqm_preset = preset;
    h->param.rc.i_qp_min = 0;
    h->param.rc.i_qp_max = QP_MAX_SPEC;
    x264_cqm_init( h );                      /* like checkasm, the "impossible QP constraints" verdict is not of interest here */
    x264_quant_init( h, 0, &g_qf );
    return 0;
}

/* quant.c:75-86 */
EXPORT int ref_quant_4x4x4_batch( int16_t *coefs, int64_t ngroups, uint16_t *mf, uint16_t *bias, int32_t *nza )
{
    ensure_tables();
    for( int64_t i = 0; i < ngroups; i++ )
        nza[i] = g_qf.quant_4x4x4( (dctcoef(*)[16])( coefs + i*64 ), mf, bias );
    return 0;
}

/* x264_mc_functions_t.weight[widx] (mc.c:119-137, widths 4,8,12,16,20 = widx 1..5 as checkasm.c:1358 numbers them) */
EXPORT void ref_mc_weight( uint8_t *dst, intptr_t ds, uint8_t *src, intptr_t ss, int widx, int height, int scale, int denom, int offset )
{
    ensure_tables();
    x264_weight_t w;
    memset( &w, 0, sizeof(w) );
    w.i_scale = scale; w.i_denom = denom; w.i_offset = offset;
    g_mc.weight[widx]( dst, ds, src, ss, &w, height );
}

EXPORT void ref_lowres_core( uint8_t *src, intptr_t src_stride, uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc,
                             intptr_t dst_stride, int width, int height )
{
    ensure_tables();
    g_mc.frame_init_lowres_core( src, d0, dh, dv, dc, src_stride, dst_stride, width, height );
}

/* mc_luma semantics (always writes dst); planes[4] as in mc.c:198; weight: denom<0 => none */
EXPORT void ref_mc_luma( uint8_t *dst, intptr_t dst_stride, uint8_t *p0, uint8_t *p1, uint8_t *p2, uint8_t *p3,
                         intptr_t src_stride, int mvx, int mvy, int w, int h, int scale, int denom, int offset )
{
    ensure_tables();
    uint8_t *src[4] = { p0, p1, p2, p3 };
    x264_weight_t wt[3];
    memset( wt, 0, sizeof(wt) );
    if( denom >= 0 )
    {
        wt[0].i_scale = scale; wt[0].i_denom = denom; wt[0].i_offset = offset;
        wt[0].weightfn = g_mc.weight;
    }
    g_mc.mc_luma( dst, dst_stride, src, src_stride, mvx, mvy, w, h, wt );
}

EXPORT void ref_avg( int size, uint8_t *dst, intptr_t ds, uint8_t *a, intptr_t as, uint8_t *b, intptr_t bs, int weight )
{
    ensure_tables();
    g_mc.avg[size]( dst, ds, a, as, b, bs, weight );
}

/* ------------------------------------------------------------------------------------------ */
/* lookahead                                                                                   */
/* ------------------------------------------------------------------------------------------ */
#define LA_MAX_FRAMES 512
typedef struct
{
    x264_t *h;
    x264_mb_analysis_t a;
    x264_frame_t *frames[LA_MAX_FRAMES + 2];
    int n_frames;
    int16_t (*snap[2])[2];
    int16_t (*newmv[2])[2];
} ref_la_t;

EXPORT void *ref_la_open( int width, int height, int bframes, int subme, int mbtree, int aq_mode,
                          int weightp, int b_adapt, int la_threads, int rc_lookahead, int vbv )
{
    ensure_tables();
    ref_la_t *c = calloc( 1, sizeof(*c) );
    if( !c ) return NULL;
    c->h = open_encoder( width, height, bframes, subme, mbtree, aq_mode, weightp, b_adapt, la_threads, rc_lookahead, vbv );
    if( !c->h ) { free( c ); return NULL; }
    int n = c->h->mb.i_mb_count;
    for( int l = 0; l < 2; l++ )
    {
        c->snap[l]  = calloc( n, 4 );
        c->newmv[l] = calloc( n, 4 );
    }
    lowres_context_init( c->h, &c->a );
    return c;
}

EXPORT void ref_la_close( void *ctx )
{
    ref_la_t *c = ctx;
    if( !c ) return;
    for( int i = 0; i < LA_MAX_FRAMES; i++ )
        if( c->frames[i] ) x264_frame_push_unused( c->h, c->frames[i] );
    x264_encoder_close( c->h );
    for( int l = 0; l < 2; l++ ) { free( c->snap[l] ); free( c->newmv[l] ); }
    free( c );
}

EXPORT int ref_la_info( void *ctx, int *out )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    out[0] = h->mb.i_mb_width; out[1] = h->mb.i_mb_height;
    out[2] = h->param.analyse.i_mv_range; out[3] = h->param.analyse.i_me_range;
    out[4] = h->mb.i_me_method; out[5] = h->mb.i_subpel_refine;
    out[6] = c->a.i_lambda; out[7] = h->param.i_bframe_bias;
    out[8] = h->param.analyse.b_weighted_bipred;
    out[9] = h->param.rc.b_mb_tree || h->param.rc.i_vbv_buffer_size || h->mb.i_mb_width <= 2 || h->mb.i_mb_height <= 2;
    return h->param.i_lookahead_threads;
}

/* copy of the lookahead mv-cost table centre +-n (analyse.c:143-157), qp = X264_LOOKAHEAD_QP */
EXPORT int ref_la_cost_mv( void *ctx, uint16_t *out, int n )
{
    ref_la_t *c = ctx;
    for( int i = -n; i <= n; i++ )
        out[i+n] = c->a.p_cost_mv[i];
    return 0;
}

/* Put frame `idx`: luma is width x height (unpadded), copied into a real x264_frame_t, padded to
 * mod16 (frame.c:640) and downscaled with x264_frame_init_lowres (mc.c:458). inv_qscale: NULL -> 256. */
EXPORT int ref_la_put_frame( void *ctx, int idx, const uint8_t *luma, int stride, const uint16_t *inv_qscale )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    if( idx < 0 || idx >= LA_MAX_FRAMES ) return -1;
    if( c->frames[idx] ) x264_frame_push_unused( h, c->frames[idx] );
    x264_frame_t *f = x264_frame_pop_unused( h, 0 );
    if( !f ) return -1;
    c->frames[idx] = f;
    f->i_frame = idx;
    f->i_type = X264_TYPE_AUTO;
    for( int y = 0; y < h->param.i_height; y++ )
        memcpy( f->plane[0] + (intptr_t)y*f->i_stride[0], luma + (intptr_t)y*stride, h->param.i_width );
    /* chroma is irrelevant to the lookahead (b_chroma_me = 0, slicetype.c:60) */
    if( h->param.i_width != 16*h->mb.i_mb_width || h->param.i_height != 16*h->mb.i_mb_height )
        x264_frame_expand_border_mod16( h, f );
    if( h->param.rc.i_aq_mode )
        for( int i = 0; i < h->mb.i_mb_count; i++ )
            f->i_inv_qscale_factor[i] = inv_qscale ? inv_qscale[i] : 256;
    x264_frame_init_lowres( h, f );
    /* x264_frame_pop_unused recycles: make never-evaluated border blocks deterministic */
    for( int j = 0; j <= !!h->param.i_bframe; j++ )
        for( int i = 0; i <= h->param.i_bframe; i++ )
        {
            memset( f->lowres_mvs[j][i], 0, 2*h->mb.i_mb_count*sizeof(int16_t) );
            f->lowres_mvs[j][i][0][0] = 0x7FFF;
            memset( f->lowres_mv_costs[j][i], 0, h->mb.i_mb_count*sizeof(int) );
        }
    for( int j = 0; j <= h->param.i_bframe+1; j++ )
        for( int i = 0; i <= h->param.i_bframe+1; i++ )
            memset( f->lowres_costs[j][i], (i|j) ? 0 : -1, h->mb.i_mb_count*sizeof(uint16_t) );
    f->b_intra_calculated = 0;
    if( idx >= c->n_frames ) c->n_frames = idx+1;
    return 0;
}

EXPORT int ref_la_release( void *ctx, int idx )
{
    ref_la_t *c = ctx;
    if( idx < 0 || idx >= LA_MAX_FRAMES || !c->frames[idx] ) return -1;
    x264_frame_push_unused( c->h, c->frames[idx] );
    c->frames[idx] = NULL;
    return 0;
}

EXPORT int ref_la_lowres_geometry( void *ctx, int idx, int *out )
{
    ref_la_t *c = ctx;
    x264_frame_t *f = c->frames[idx];
    out[0] = f->i_width_lowres; out[1] = f->i_lines_lowres; out[2] = f->i_stride_lowres;
    out[3] = PADH; out[4] = PADV;
    return 0;
}

/* copies plane `p` including its PADH/PADV border into out (row pitch = width+2*PADH) */
EXPORT int ref_la_get_lowres( void *ctx, int idx, int p, uint8_t *out )
{
    ref_la_t *c = ctx;
    x264_frame_t *f = c->frames[idx];
    int w = f->i_width_lowres + 2*PADH;
    for( int y = -PADV; y < f->i_lines_lowres + PADV; y++ )
        memcpy( out + (intptr_t)(y+PADV)*w, f->lowres[p] + (intptr_t)y*f->i_stride_lowres - PADH, w );
    return 0;
}

/* The reference's own raster-serial evaluation (slicetype.c:836). */
EXPORT int ref_la_frame_cost_serial( void *ctx, int p0, int p1, int b )
{
    ref_la_t *c = ctx;
    return slicetype_frame_cost( c->h, &c->a, c->frames, p0, p1, b );
}

/* The dependency-free schedule.  Pass k evaluates every block with the UNMODIFIED slicetype_mb_cost,
 * after restoring the four neighbour MVs it reads (slicetype.c:666-674: right, below, below-left,
 * below-right) to their values from pass k-1 (pass 0: the zero field).  Accumulation and the
 * epilogue follow slicetype.c:946-989 for a single lookahead slice.  weights: not analysed here
 * (w = x264_weight_none) unless a weight is installed with ref_la_set_weight. */
EXPORT int ref_la_frame_cost_jacobi( void *ctx, int p0, int p1, int b, int passes )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    x264_frame_t **frames = c->frames;
    x264_frame_t *fenc = frames[b];
    const x264_weight_t *w = x264_weight_none;
    int do_search[2];
    int mbw = h->mb.i_mb_width, mbh = h->mb.i_mb_height, n = h->mb.i_mb_count;

    if( fenc->i_cost_est[b-p0][p1-b] >= 0 && (!h->param.rc.i_vbv_buffer_size || fenc->i_row_satds[b-p0][p1-b][0] != -1) )
        return fenc->i_cost_est[b-p0][p1-b];

    int dist_scale_factor = 128;
    do_search[0] = b != p0 && fenc->lowres_mvs[0][b-p0-1][0][0] == 0x7FFF;
    do_search[1] = b != p1 && fenc->lowres_mvs[1][p1-b-1][0][0] == 0x7FFF;
    if( do_search[0] )
    {
        if( h->param.analyse.i_weighted_pred && b == p1 )
        {
            x264_weights_analyse( h, fenc, frames[p0], 1 );
            w = fenc->weight[0];
        }
        fenc->lowres_mvs[0][b-p0-1][0][0] = 0;
    }
    if( do_search[1] ) fenc->lowres_mvs[1][p1-b-1][0][0] = 0;
    if( p1 != p0 )
        dist_scale_factor = ( ((b-p0) << 8) + ((p1-p0) >> 1) ) / (p1-p0);

    int16_t (*live[2])[2] = { b != p0 ? fenc->lowres_mvs[0][b-p0-1] : NULL, b != p1 ? fenc->lowres_mvs[1][p1-b-1] : NULL };
    int output_buf_size = mbh + (NUM_INTS + PAD_SIZE);
    int *output_inter = h->scratch_buffer2;
    int *output_intra = output_inter + output_buf_size;

    int do_edges = h->param.rc.b_mb_tree || h->param.rc.i_vbv_buffer_size || mbw <= 2 || mbh <= 2;
    h->i_threadslice_start = 0;
    h->i_threadslice_end = mbh;
    int start_y = X264_MIN( mbh - 1, mbh - 2 + do_edges );
    int end_y = X264_MAX( 0, 1 - do_edges );
    int start_x = mbw - 2 + do_edges;
    int end_x = 1 - do_edges;
    int n_pass = (do_search[0] || do_search[1]) ? passes : 1;

    for( int l = 0; l < 2; l++ )
        if( do_search[l] )
            memset( live[l], 0, n*4 );

    for( int pass = 0; pass < n_pass; pass++ )
    {
        for( int l = 0; l < 2; l++ )
            if( do_search[l] )
            {
                memcpy( c->snap[l], live[l], n*4 );
                memcpy( c->newmv[l], live[l], n*4 );
            }
        memset( output_inter, 0, (output_buf_size - PAD_SIZE) * sizeof(int) );
        memset( output_intra, 0, (output_buf_size - PAD_SIZE) * sizeof(int) );
        output_inter[NUM_ROWS] = output_intra[NUM_ROWS] = mbh;

        for( h->mb.i_mb_y = start_y; h->mb.i_mb_y >= end_y; h->mb.i_mb_y-- )
            for( h->mb.i_mb_x = start_x; h->mb.i_mb_x >= end_x; h->mb.i_mb_x-- )
            {
                int x = h->mb.i_mb_x, y = h->mb.i_mb_y, xy = x + y*mbw;
                for( int l = 0; l < 2; l++ )
                    if( do_search[l] )
                    {
                        if( x < mbw-1 ) CP32( live[l][xy+1], c->snap[l][xy+1] );
                        if( y < mbh-1 )
                        {
                            CP32( live[l][xy+mbw], c->snap[l][xy+mbw] );
                            if( x > 0 )     CP32( live[l][xy+mbw-1], c->snap[l][xy+mbw-1] );
                            if( x < mbw-1 ) CP32( live[l][xy+mbw+1], c->snap[l][xy+mbw+1] );
                        }
                    }
                slicetype_mb_cost( h, &c->a, frames, p0, p1, b, dist_scale_factor, do_search, w, output_inter, output_intra );
                for( int l = 0; l < 2; l++ )
                    if( do_search[l] )
                        CP32( c->newmv[l][xy], live[l][xy] );
            }
        for( int l = 0; l < 2; l++ )
            if( do_search[l] )
                memcpy( live[l], c->newmv[l], n*4 );
    }

    /* epilogue: slicetype.c:946-989 with i_lookahead_threads == 1 (same statement order: for an
     * I evaluation [b-p0][p1-b] aliases [0][0]) */
    if( b == p1 )
        fenc->i_intra_mbs[b-p0] = 0;
    if( !fenc->b_intra_calculated )
    {
        fenc->i_cost_est[0][0] = 0;
        fenc->i_cost_est_aq[0][0] = 0;
    }
    fenc->i_cost_est[b-p0][p1-b] = 0;
    fenc->i_cost_est_aq[b-p0][p1-b] = 0;
    if( b == p1 )
        fenc->i_intra_mbs[b-p0] += output_inter[INTRA_MBS];
    if( !fenc->b_intra_calculated )
    {
        fenc->i_cost_est[0][0] += output_intra[COST_EST];
        fenc->i_cost_est_aq[0][0] += output_intra[COST_EST_AQ];
    }
    fenc->i_cost_est[b-p0][p1-b] += output_inter[COST_EST];
    fenc->i_cost_est_aq[b-p0][p1-b] += output_inter[COST_EST_AQ];
    /* row SATDs are always exported by this oracle (the reference copies them only under VBV) */
    memcpy( fenc->i_row_satds[b-p0][p1-b], output_inter + NUM_INTS, mbh * sizeof(int) );
    if( !fenc->b_intra_calculated )
        memcpy( fenc->i_row_satds[0][0], output_intra + NUM_INTS, mbh * sizeof(int) );

    int i_score = fenc->i_cost_est[b-p0][p1-b];
    if( b != p1 )
        i_score = (uint64_t)i_score * 100 / (120 + h->param.i_bframe_bias);
    else
        fenc->b_intra_calculated = 1;
    fenc->i_cost_est[b-p0][p1-b] = i_score;
    return i_score;
}

/* what: 0 lowres_costs[i][j] (u16 x n), 1 lowres_mvs[i][j] (i16 x 2n), 2 lowres_mv_costs[i][j] (i32 x n),
 *       3 i_row_satds[i][j] (i32 x mbh), 4 {cost_est[i][j], cost_est_aq[i][j], intra_mbs[i]} (i32 x 3),
 *       5 i_intra_cost (u16 x n), 6 i_inv_qscale_factor (u16 x n) */
EXPORT int ref_la_get( void *ctx, int idx, int what, int i, int j, void *out )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    x264_frame_t *f = c->frames[idx];
    int n = h->mb.i_mb_count;
    switch( what )
    {
        case 0: memcpy( out, f->lowres_costs[i][j], n*2 ); break;
        case 1: memcpy( out, f->lowres_mvs[i][j], n*4 ); break;
        case 2: memcpy( out, f->lowres_mv_costs[i][j], n*4 ); break;
        case 3: memcpy( out, f->i_row_satds[i][j], h->mb.i_mb_height*4 ); break;
        case 4: ((int*)out)[0] = f->i_cost_est[i][j]; ((int*)out)[1] = f->i_cost_est_aq[i][j]; ((int*)out)[2] = f->i_intra_mbs[i]; break;
        case 5: memcpy( out, f->i_intra_cost, n*2 ); break;
        case 6: memcpy( out, f->i_inv_qscale_factor, n*2 ); break;
        default: return -1;
    }
    return 0;
}

/* Real AQ producer for inv_qscale (ratecontrol.c:304), so tests can feed realistic values. */
EXPORT int ref_la_adaptive_quant( void *ctx, int idx )
{
    ref_la_t *c = ctx;
    x264_adaptive_quant_frame( c->h, c->frames[idx], NULL );
    return 0;
}

/* Run the reference's own slicetype analysis over frames[0..n) as a lookahead window:
 * frames[0] is the last non-B frame, the rest are undecided.  Returns decided types. */
EXPORT int ref_la_slicetype_analyse( void *ctx, int first, int n, int intra_minigop, uint8_t *types_out )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    if( n < 2 || n - 1 > h->lookahead->next.i_max_size ) return -1;
    h->lookahead->last_nonb = c->frames[first];
    c->frames[first]->i_type = X264_TYPE_I;
    h->lookahead->next.i_size = 0;
    for( int i = 1; i < n; i++ )
    {
        c->frames[first+i]->i_type = X264_TYPE_AUTO;
        h->lookahead->next.list[h->lookahead->next.i_size++] = c->frames[first+i];
    }
    h->lookahead->next.list[h->lookahead->next.i_size] = NULL;
    x264_slicetype_analyse( h, intra_minigop );
    for( int i = 0; i < n; i++ )
        types_out[i] = c->frames[first+i]->i_type;
    h->lookahead->next.i_size = 0;
    h->lookahead->next.list[0] = NULL;
    h->lookahead->last_nonb = NULL;       /* x264_lookahead_delete would free it (lookahead.c) */
    return 0;
}

/* Pre-populate the cost caches of frame idx so that the reference's own decision logic runs on externally
 * supplied costs (every slicetype_frame_cost becomes a cache hit, slicetype.c:848). */
EXPORT int ref_la_inject_cost( void *ctx, int idx, int d0, int d1, int cost_est, int cost_est_aq, int intra_mbs )
{
    ref_la_t *c = ctx;
    x264_frame_t *f = c->frames[idx];
    if( !f ) return -1;
    f->i_cost_est[d0][d1] = cost_est;
    f->i_cost_est_aq[d0][d1] = cost_est_aq;
    if( d1 == 0 ) f->i_intra_mbs[d0] = intra_mbs;
    f->i_row_satds[d0][d1][0] = 0;
    return 0;
}

/* x264_slicetype_analyse (slicetype.c:1473) over frames[first..first+n): frames[first] is the last non-B frame.
 * forced[i] = forced type of frames[first+i] (0 = auto).  Writes the types after the call and the frames' i_type
 * as decided; returns 0. */
EXPORT int ref_la_analyse2( void *ctx, int first, int n, int intra_minigop, int last_keyframe, const uint8_t *forced, uint8_t *types_out )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    if( n < 1 || n - 1 > h->lookahead->next.i_max_size ) return -1;   /* the real lookahead never holds more */
    h->lookahead->last_nonb = c->frames[first];
    h->lookahead->i_last_keyframe = last_keyframe;
    h->lookahead->next.i_size = 0;
    for( int i = 1; i < n; i++ )
    {
        x264_frame_t *f = c->frames[first+i];
        f->i_forced_type = forced ? forced[i] : X264_TYPE_AUTO;
        f->i_type = f->i_forced_type;
        f->b_scenecut = 1;
        h->lookahead->next.list[h->lookahead->next.i_size++] = f;
    }
    h->lookahead->next.list[h->lookahead->next.i_size] = NULL;
    x264_slicetype_analyse( h, intra_minigop );
    for( int i = 0; i < n; i++ )
        types_out[i] = c->frames[first+i]->i_type;
    h->lookahead->next.i_size = 0;
    h->lookahead->next.list[0] = NULL;
    h->lookahead->last_nonb = NULL;       /* x264_lookahead_delete would free it (lookahead.c) */
    return 0;
}

EXPORT int ref_la_param( void *ctx, int *out )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    out[0] = h->param.i_bframe; out[1] = h->param.i_bframe_adaptive; out[2] = h->param.i_bframe_pyramid;
    out[3] = h->param.i_scenecut_threshold; out[4] = h->param.i_keyint_max; out[5] = h->param.i_keyint_min;
    out[6] = h->param.b_open_gop; out[7] = h->param.b_intra_refresh;
    out[8] = (h->param.analyse.b_psy && h->param.rc.b_mb_tree) || (h->param.rc.i_vbv_buffer_size && h->param.rc.i_lookahead);
    out[9] = h->lookahead->i_slicetype_length; out[10] = h->param.b_deterministic;
    return 0;
}

/* ---- MB-tree ---- */
EXPORT void ref_mbtree_propagate_cost( int16_t *dst, uint16_t *propagate_in, uint16_t *intra_costs, uint16_t *inter_costs,
                                       uint16_t *inv_qscales, float fps, int len )
{
    ensure_tables();
    g_mc.mbtree_propagate_cost( dst, propagate_in, intra_costs, inter_costs, inv_qscales, &fps, len );
}
EXPORT void ref_mbtree_propagate_list( void *ctx, uint16_t *ref_costs, int16_t *mvs, int16_t *propagate_amount, uint16_t *lowres_costs,
                                       int bipred_weight, int mb_y, int len, int list )
{
    ref_la_t *c = ctx;
    ensure_tables();
    g_mc.mbtree_propagate_list( c->h, ref_costs, (int16_t(*)[2])mvs, propagate_amount, lowres_costs, bipred_weight, mb_y, len, list );
}
EXPORT int ref_log2_lut( float *out128 ) { memcpy( out128, x264_log2_lut, 128*sizeof(float) ); return 0; }

/* like ref_la_put_frame, but runs the reference's adaptive quantisation first (x264_adaptive_quant_frame,
 * ratecontrol.c:304, as encoder.c:3417 does before x264_frame_init_lowres): fills f_qp_offset_aq and
 * i_inv_qscale_factor from the picture itself */
EXPORT int ref_la_put_frame_aq( void *ctx, int idx, const uint8_t *luma, int stride )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    if( ref_la_put_frame( ctx, idx, luma, stride, NULL ) ) return -1;
    x264_frame_t *f = c->frames[idx];
    for( int y = 0; y < f->i_lines[1]; y++ )
        memset( f->plane[1] + (intptr_t)y*f->i_stride[1], 128, f->i_width[1] );
    x264_adaptive_quant_frame( h, f, NULL );
    return 0;
}

/* what: 0 f_qp_offset_aq, 1 f_qp_offset (float x n), 2 i_propagate_cost (u16 x n) */
EXPORT int ref_la_get_mbtree( void *ctx, int idx, int what, void *out )
{
    ref_la_t *c = ctx;
    x264_frame_t *f = c->frames[idx];
    int n = c->h->mb.i_mb_count;
    if( what == 0 ) memcpy( out, f->f_qp_offset_aq, n*sizeof(float) );
    else if( what == 1 ) memcpy( out, f->f_qp_offset, n*sizeof(float) );
    else if( what == 2 ) memcpy( out, f->i_propagate_cost, n*sizeof(uint16_t) );
    else return -1;
    return 0;
}

/* the reference's own macroblock_tree (slicetype.c:1091) over frames[first..first+num_frames]; every
 * slicetype_frame_cost inside it is a cache hit when the caller has evaluated the triples beforehand */
EXPORT int ref_la_macroblock_tree( void *ctx, int first, int num_frames, int b_intra, const uint8_t *types, const float *durations )
{
    ref_la_t *c = ctx;
    for( int i = 0; i <= num_frames; i++ )
    {
        c->frames[first+i]->i_type = types[i];
        c->frames[first+i]->f_duration = durations[i];
        memset( c->frames[first+i]->f_weighted_cost_delta, 0, sizeof(c->frames[first+i]->f_weighted_cost_delta) );
    }
    macroblock_tree( c->h, &c->a, c->frames + first, num_frames, b_intra );
    return 0;
}

/* ---- adaptive quantisation + weighted-prediction analysis ---- */
/* like ref_la_put_frame_aq with the real interleaved NV12 chroma plane: x264_adaptive_quant_frame (ratecontrol.c:304)
 * needs both chroma components for ac_energy_mb (:262-294) */
EXPORT int ref_la_put_frame_yuv( void *ctx, int idx, const uint8_t *luma, int stride, const uint8_t *chroma, int cstride,
                                 const float *quant_offsets )
{
    ref_la_t *c = ctx;
    x264_t *h = c->h;
    if( ref_la_put_frame( ctx, idx, luma, stride, NULL ) ) return -1;
    x264_frame_t *f = c->frames[idx];
    for( int y = 0; y < h->param.i_height/2; y++ )
        memcpy( f->plane[1] + (intptr_t)y*f->i_stride[1], chroma + (intptr_t)y*cstride, h->param.i_width );
    if( h->param.i_width != 16*h->mb.i_mb_width || h->param.i_height != 16*h->mb.i_mb_height )
        x264_frame_expand_border_mod16( h, f );
    x264_adaptive_quant_frame( h, f, (float *)quant_offsets );
    return 0;
}

/* f_qp_offset, i_inv_qscale_factor, {i_pixel_sum[p], i_pixel_ssd[p]} p = 0..2 */
EXPORT int ref_la_get_aq( void *ctx, int idx, float *qp_offset, uint16_t *inv_qscale, uint64_t *sums6 )
{
    ref_la_t *c = ctx;
    x264_frame_t *f = c->frames[idx];
    int n = c->h->mb.i_mb_count;
    if( qp_offset ) memcpy( qp_offset, f->f_qp_offset, n*sizeof(float) );
    if( inv_qscale ) memcpy( inv_qscale, f->i_inv_qscale_factor, n*sizeof(uint16_t) );
    if( sums6 ) for( int i = 0; i < 3; i++ ) { sums6[2*i] = f->i_pixel_sum[i]; sums6[2*i+1] = f->i_pixel_ssd[i]; }
    return 0;
}

/* the reference's own x264_weights_analyse( h, frames[b], frames[ref], b_lookahead = 1 ) (slicetype.c:284);
 * out: [0] weightfn != NULL [1] i_scale [2] i_denom [3] i_offset */
EXPORT int ref_la_weights_analyse( void *ctx, int b, int ref, int *out )
{
    ref_la_t *c = ctx;
    x264_frame_t *fenc = c->frames[b];
    x264_weights_analyse( c->h, fenc, c->frames[ref], 1 );
    out[0] = fenc->weight[0][0].weightfn != NULL;
    out[1] = fenc->weight[0][0].i_scale; out[2] = fenc->weight[0][0].i_denom; out[3] = fenc->weight[0][0].i_offset;
    return 0;
}

EXPORT int ref_exp2_lut( uint8_t *out64 ) { memcpy( out64, x264_exp2_lut, 64 ); return 0; }

EXPORT double ref_now( void )
{
    struct timespec ts;
    clock_gettime( CLOCK_MONOTONIC, &ts );
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}
