/* x264_glue.c -- the reference-side binding: implements the entry points the reference declares in
 * encoder/slicetype-cl.h:29-42 and common/opencl.h:796-807 on top of libx264cuda.so, so that
 * common/opencl.c + encoder/slicetype-cl.c can be dropped from the build (Makefile:240-245) and every
 * hook site in slicetype.c / encoder.c / frame.c stays as it is.
 *
 * This file is compiled INSIDE an x264 tree (it includes x264's private headers); it is not part of
 * libx264cuda.so.  oracle/Makefile's `glue-check` target type-checks it against /root/reference, and its `cli`
 * target links it with the reference's unmodified encoder + CLI sources into oracle/_ref/x264cu_cli, which
 * tests/test_gpu_x264_cli.py runs on a B200: `x264 --opencl` then encodes through this backend, and in the exact
 * search mode the slices it writes are byte-identical to the CPU-only run.
 *
 * State: the reference embeds x264_opencl_t in x264_t by value and copies x264_t into the lookahead
 * thread (encoder/lookahead.c:153-154), so everything mutable lives behind ONE heap pointer that we
 * park in h->opencl.context; the per-frame handle is parked in
 * frame->opencl.luma_hpel.
 */
#include "common/common.h"
#include "encoder/slicetype-cl.h"
#include "x264cu.h"

/* declared privately by the reference in slicetype.c:38-39 and slicetype-cl.c:29-30 */
#define x264_weights_analyse x264_template(weights_analyse)
void x264_weights_analyse( x264_t *h, x264_frame_t *fenc, x264_frame_t *ref, int b_lookahead );

#if HAVE_OPENCL   /* the reference's own gate (configure:1425-1428): 8-bit builds only */

/* The scalar results of an evaluation (i_cost_est & co.) are ints scattered over x264_frame_t; the device delivers
 * them into a staging slot when the queue is flushed, and the flush copies them to their fields -- the same
 * deferred-copy protocol as the reference's (slicetype-cl.c:58-70). */
typedef struct
{
    x264_frame_t *f;
    int d0, d1;                 /* b-p0, p1-b; 0,0: the intra results of x264_opencl_lowres_init */
    int32_t st[8];              /* x264cu_out.stats (4) + intra_stats (2) */
} cu_pending;

#define CU_MAX_PENDING 1024     /* the reference's own bound on queued copies (opencl.h:715) */
typedef struct
{
    x264cu_ctx *c;
    int n_pend;
    cu_pending pend[CU_MAX_PENDING];
} cu_state;

#define CU_STATE(h) ((cu_state *)(h)->opencl.context)
#define CU_CTX(h)   (CU_STATE(h)->c)
#define CU_FRAME(f) ((x264cu_frame *)(f)->opencl.luma_hpel)

static int cu_fail( x264_t *h, const char *what )
{   /* OCLCHECK convention (slicetype-cl.c:44-56); no CPU fallback: the error is fatal */
    x264_log( h, X264_LOG_ERROR, "x264cu: %s failed: %s\n", what, x264cu_last_error() );
    h->param.b_opencl = 0;
    h->opencl.b_fatal_error = 1;
    return -1;
}

static cu_pending *cu_pend( x264_t *h, x264_frame_t *f, int d0, int d1 )
{
    cu_state *s = CU_STATE( h );
    if( s->n_pend == CU_MAX_PENDING )
        x264_opencl_flush( h );                          /* slots are handed to the device by address */
    cu_pending *p = &s->pend[s->n_pend++];
    memset( p, 0, sizeof(*p) );
    p->f = f; p->d0 = d0; p->d1 = d1;
    return p;
}

/* opencl.h:796-799.  There is no library to dlopen; hand back a non-NULL token. */
x264_opencl_function_t *x264_opencl_load_library( void )
{
    static x264_opencl_function_t token;
    return x264cu_device_count() > 0 ? &token : NULL;
}
void x264_opencl_close_library( x264_opencl_function_t *ocl ) { (void)ocl; }

/* opencl.h:801-802, called from encoder.c:1797-1800 before x264_lookahead_init copies *h */
int x264_opencl_lookahead_init( x264_t *h )
{
    x264cu_cfg cfg;
    memset( &cfg, 0, sizeof(cfg) );
    cfg.width  = h->param.i_width;
    cfg.height = h->param.i_height;
    cfg.bframes = h->param.i_bframe;
    cfg.subme = h->param.analyse.i_subpel_refine;
    cfg.me_method = h->param.analyse.i_me_method;
    cfg.me_range = h->param.analyse.i_me_range;
    cfg.mv_range = h->param.analyse.i_mv_range;
    cfg.do_edges = h->param.rc.b_mb_tree || h->param.rc.i_vbv_buffer_size;
    cfg.aq = h->param.rc.i_aq_mode != 0;
    cfg.weighted_bipred = h->param.analyse.b_weighted_bipred;
    cfg.bframe_bias = h->param.i_bframe_bias;
    /* 2: the dependency-free schedule; 0: the reference's own raster-serial order, whose
     * costs, MVs and therefore decisions and bitstream equal the CPU path's.  Selected by the encoder's own parameters:
     * with ONE lookahead thread (--lookahead-threads 1, or --threads 1) the CPU lookahead is a single raster-serial pass
     * over the frame (slicetype.c:901-944 runs no slices), and a deterministic encode (b_deterministic, the default) is
     * then reproduced bit for bit by the exact mode; with several lookahead threads the CPU path cuts the frame into
     * slices of its own and the throughput-oriented schedule is used.  X264CU_SEARCH_PASSES overrides (tests, tuning). */
    const char *e = getenv( "X264CU_SEARCH_PASSES" );
    cfg.search_passes = e ? atoi( e ) : ( h->param.i_lookahead_threads == 1 && h->param.b_deterministic ) ? 0 : 2;
    cfg.max_frames = h->param.rc.i_lookahead + h->param.i_bframe + 3;
    cfg.cost_mv = h->cost_mv[X264_LOOKAHEAD_QP];          /* the encoder's own table (analyse.c:143-157) */
    cu_state *s = calloc( 1, sizeof(*s) );
    if( !s )
        return -1;
    s->c = x264cu_open( h->param.i_opencl_device, &cfg );
    if( !s->c )
    {
        x264_log( h, X264_LOG_ERROR, "x264cu: %s\n", x264cu_last_error() );
        free( s );
        return -1;
    }
    h->opencl.context = (cl_context)s;
    h->opencl.b_fatal_error = 0;
    return 0;
}

/* opencl.h:803-804, called from encoder.c:4207-4209,4574-4576 */
void x264_opencl_lookahead_delete( x264_t *h )
{
    cu_state *s = CU_STATE( h );
    if( !s )
        return;
    x264cu_close( s->c );
    free( s );
    h->opencl.context = NULL;
}

/* opencl.h:806-807, called from frame.c:335-337 */
void x264_opencl_frame_delete( x264_frame_t *frame )
{
    /* the context owns its frames and frees the survivors in x264cu_close; frames recycled through
     * x264_frame_pop_unused keep their device objects exactly like the reference (frame.c:780-801) */
    frame->opencl.luma_hpel = NULL;
}

/* slicetype-cl.h:29-30: x264_frame_init_lowres on the device + intra costs, once per frame; the intra results
 * (i_intra_cost, i_row_satds[0][0], i_cost_est[0][0], i_cost_est_aq[0][0]) reach the frame with the next flush,
 * as in the reference (slicetype-cl.c:254-280) */
int x264_opencl_lowres_init( x264_t *h, x264_frame_t *fenc, int lambda )
{
    (void)lambda;                                        /* lambda == 1 at X264_LOOKAHEAD_QP */
    if( fenc->b_intra_calculated && CU_FRAME( fenc ) )
        return 0;
    if( !CU_FRAME( fenc ) )
    {
        x264cu_frame *f = x264cu_frame_new( CU_CTX( h ) );
        if( !f ) return cu_fail( h, "x264cu_frame_new" );
        fenc->opencl.luma_hpel = (cl_mem)f;
    }
    /* same bytes the reference uploads (slicetype-cl.c:89,166-177): the luma plane and inv_qscale */
    if( x264cu_frame_upload( CU_CTX( h ), CU_FRAME( fenc ), fenc->plane[0], fenc->i_stride[0],
                             h->param.rc.i_aq_mode ? fenc->i_inv_qscale_factor : NULL ) )
        return cu_fail( h, "x264cu_frame_upload" );
#ifdef X264CU_GLUE_DEVICE_WEIGHTS
    /* x264_adaptive_quant_frame ran on the host (encoder.c:3417): give its luma statistics to the device so that
     * x264cu_weights_analyse can replace x264_weights_analyse without a host round trip */
    if( h->param.analyse.i_weighted_pred &&
        x264cu_frame_set_pixel_stats( CU_CTX( h ), CU_FRAME( fenc ), fenc->i_pixel_sum[0], fenc->i_pixel_ssd[0] ) )
        return cu_fail( h, "x264cu_frame_set_pixel_stats" );
#endif
    cu_pending *p = cu_pend( h, fenc, 0, 0 );
    x264cu_job j;
    memset( &j, 0, sizeof(j) );
    j.fenc = j.fref0 = j.fref1 = CU_FRAME( fenc );
    j.out.intra_costs = fenc->i_intra_cost;
    j.out.intra_row_satds = fenc->i_row_satds[0][0];
    j.out.intra_stats = p->st + 4;
    if( x264cu_frame_cost( CU_CTX( h ), &j ) )
        return cu_fail( h, "x264cu_frame_cost" );
    fenc->b_intra_calculated = 1;                        /* set eagerly, as slicetype-cl.c:84-86 does */
    return 0;
}

/* slicetype-cl.h:31-32 */
int x264_opencl_motionsearch( x264_t *h, x264_frame_t **frames, int b, int ref, int b_islist1, int lambda, const x264_weight_t *w )
{
    (void)lambda;
    x264cu_weight wt = { 0, 0, 0, 0 };
    if( w && w->weightfn ) { wt.enabled = 1; wt.scale = w->i_scale; wt.denom = w->i_denom; wt.offset = w->i_offset; }
    int dist = b_islist1 ? ref - b : b - ref;
    if( x264cu_motionsearch( CU_CTX( h ), CU_FRAME( frames[b] ), CU_FRAME( frames[ref] ), b_islist1, dist, &wt,
                             (int16_t *)frames[b]->lowres_mvs[b_islist1][dist-1],
                             frames[b]->lowres_mv_costs[b_islist1][dist-1] ) )
        return cu_fail( h, "x264cu_motionsearch" );
    return 0;
}

/* slicetype-cl.h:33-34: mode decision + sums; results land in the x264_frame_t arrays at flush */
int x264_opencl_finalize_cost( x264_t *h, int lambda, x264_frame_t **frames, int p0, int p1, int b, int dist_scale_factor )
{
    (void)lambda; (void)dist_scale_factor;               /* recomputed from (b-p0, p1-b), slicetype.c:869-870 */
    x264_frame_t *fenc = frames[b];
    cu_pending *p = cu_pend( h, fenc, b - p0, p1 - b );
    x264cu_job j;
    memset( &j, 0, sizeof(j) );
    j.fenc = CU_FRAME( fenc ); j.fref0 = CU_FRAME( frames[p0] ); j.fref1 = CU_FRAME( frames[p1] );
    j.d0 = b - p0; j.d1 = p1 - b;
    j.out.lowres_costs = fenc->lowres_costs[b-p0][p1-b];
    j.out.row_satds = fenc->i_row_satds[b-p0][p1-b];
    j.out.stats = p->st;
    if( x264cu_frame_cost( CU_CTX( h ), &j ) )
        return cu_fail( h, "x264cu_frame_cost" );
    return 0;
}

/* slicetype-cl.h:35-36: defined but never called by the reference */
int x264_opencl_precalculate_frame_cost( x264_t *h, x264_frame_t **frames, int lambda, int p0, int p1, int b )
{
    (void)h; (void)frames; (void)lambda; (void)p0; (void)p1; (void)b;
    return 0;
}

/* slicetype-cl.h:37-38: wait for the device, then finish the scalar results (slicetype-cl.c:58-70) */
void x264_opencl_flush( x264_t *h )
{
    cu_state *s = CU_STATE( h );
    if( x264cu_flush( s->c ) )
    {
        cu_fail( h, "x264cu_flush" );
        s->n_pend = 0;
        return;
    }
    for( int i = 0; i < s->n_pend; i++ )
    {
        cu_pending *p = &s->pend[i];
        if( !( p->d0 | p->d1 ) )
        {   /* slicetype.c:949-953: the intra sums are the I cost */
            p->f->i_cost_est[0][0] = p->st[4];
            p->f->i_cost_est_aq[0][0] = p->st[5];
        }
        else
        {   /* slicetype.c:965-989 */
            p->f->i_cost_est[p->d0][p->d1] = p->st[0];
            p->f->i_cost_est_aq[p->d0][p->d1] = p->st[1];
            if( !p->d1 ) p->f->i_intra_mbs[p->d0] = p->st[2];
        }
    }
    s->n_pend = 0;
}

/* slicetype-cl.h:39-40: intra for all frames, then (trellis) every (b, b+-j) search, one flush each */
void x264_opencl_slicetype_prep( x264_t *h, x264_frame_t **frames, int num_frames, int lambda )
{
    if( !h->param.b_opencl )
        return;
    for( int i = 0; i <= num_frames; i++ )
        x264_opencl_lowres_init( h, frames[i], lambda );
    if( h->param.i_bframe_adaptive == X264_B_ADAPT_TRELLIS && h->param.i_bframe )
        for( int b = 0; b <= num_frames; b++ )
            for( int j = 1; j < h->param.i_bframe; j++ )
            {
                int p0 = b - j, p1 = b + j;
                if( p0 >= 0 && frames[b]->lowres_mvs[0][b-p0-1][0][0] == 0x7FFF )
                {
                    const x264_weight_t *w = x264_weight_none;
#ifdef X264CU_GLUE_DEVICE_WEIGHTS
                    if( h->param.analyse.i_weighted_pred )
                    {   /* analysis, weighted plane and search all on the device; the weight reaches frames[b]->weight[0]
                         * with the flush at the end of this function (slicetype.c:284-501 without the host sync) */
                        x264cu_weight dev = { X264CU_WEIGHT_DEVICE, 0, 0, 0 };
                        frames[b]->lowres_mvs[0][b-p0-1][0][0] = 0;
                        if( x264cu_weights_analyse( CU_CTX( h ), CU_FRAME( frames[b] ), CU_FRAME( frames[p0] ), b-p0,
                                                    h->param.i_slice_count ? h->param.i_slice_count : 1, NULL ) ||
                            x264cu_motionsearch( CU_CTX( h ), CU_FRAME( frames[b] ), CU_FRAME( frames[p0] ), 0, b-p0, &dev,
                                                 (int16_t *)frames[b]->lowres_mvs[0][b-p0-1], frames[b]->lowres_mv_costs[0][b-p0-1] ) )
                            cu_fail( h, "x264cu_weights_analyse" );
                    }
                    else
#else
                    if( h->param.analyse.i_weighted_pred )
                    {
                        x264_weights_analyse( h, frames[b], frames[p0], 1 );
                        w = frames[b]->weight[0];
                    }
#endif
                    {
                    frames[b]->lowres_mvs[0][b-p0-1][0][0] = 0;
                    x264_opencl_motionsearch( h, frames, b, p0, 0, lambda, w );
                    }
                }
                if( p1 <= num_frames && frames[b]->lowres_mvs[1][p1-b-1][0][0] == 0x7FFF )
                {
                    frames[b]->lowres_mvs[1][p1-b-1][0][0] = 0;
                    x264_opencl_motionsearch( h, frames, b, p1, 1, lambda, NULL );
                }
            }
    x264_opencl_flush( h );
}

/* slicetype-cl.h:41-42 */
void x264_opencl_slicetype_end( x264_t *h ) { (void)h; }

#endif /* HAVE_OPENCL */
