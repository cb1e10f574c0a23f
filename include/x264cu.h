/* x264cu.h -- C ABI of libx264cuda.so: the B200 (sm_100a) lookahead cost engine.
 *
 * This library sits where the reference's OpenCL lookahead backend sits
 * (common/opencl.c + encoder/slicetype-cl.c behind encoder/slicetype.c:878-899,1530,1740 and
 * encoder/encoder.c:1744-1754,1797-1800).  Plain pointers and sizes only; no x264 or torch types.
 * INTEGRATION.md shows the ~200-line glue that implements the reference's x264_opencl_* entry
 * points (encoder/slicetype-cl.h:29-42, common/opencl.h:796-807) on top of these calls.
 *
 * Conventions: every int-returning call returns 0 on success and -1 on failure (the reference's
 * OCLCHECK convention, slicetype-cl.c:44-56); x264cu_last_error() describes the failure.
 * There is no CPU fallback: without a CUDA device x264cu_open fails.
 * A context is single-threaded, like the reference's one in-order queue (opencl.c:525); several
 * contexts (encode streams) may share a device, each with its own CUDA stream.
 */
#ifndef X264CU_H
#define X264CU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X264CU_BFRAME_MAX 16        /* X264_BFRAME_MAX, common/base.h:136 */

typedef struct x264cu_ctx   x264cu_ctx;     /* replaces x264_opencl_t        (common/opencl.h:720-778) */
typedef struct x264cu_frame x264cu_frame;   /* replaces x264_frame_opencl_t  (common/opencl.h:780-792) */

/* block sizes, in the order of the reference's PIXEL_* enum (common/pixel.h:41-52) */
enum { X264CU_16x16, X264CU_16x8, X264CU_8x16, X264CU_8x8, X264CU_8x4, X264CU_4x8, X264CU_4x4, X264CU_4x16 };
/* metrics of x264_pixel_function_t (common/pixel.h:82-103) */
enum { X264CU_SAD, X264CU_SATD, X264CU_SA8D };

/* The x264_param_t fields that steer the path (common/base.c:378-460). */
typedef struct
{
    int width, height;        /* i_width / i_height: luma size before mod-16 padding                      */
    int bframes;              /* i_bframe (0..16)                                                         */
    int subme;                /* analyse.i_subpel_refine: >1 selects SATD + HEX + subme 4 (slicetype.c:50) */
    int me_method;            /* analyse.i_me_method: 0 dia, >=1 hex (min(HEX,param), slicetype.c:52)      */
    int me_range;             /* analyse.i_me_range (16)                                                  */
    int mv_range;             /* analyse.i_mv_range after x264's level-based validation (e.g. 512 @1080p) */
    int do_edges;             /* rc.b_mb_tree || rc.i_vbv_buffer_size (slicetype.c:823)                   */
    int aq;                   /* rc.i_aq_mode != 0 (slicetype.c:749,779)                                  */
    int weighted_bipred;      /* analyse.b_weighted_bipred (slicetype.c:529)                              */
    int bframe_bias;          /* i_bframe_bias (slicetype.c:985)                                          */
    int search_passes;        /* K >= 1: passes of the dependency-free schedule, default 2;
                               * 0: exact raster-serial search (one launch per anti-diagonal wave): results
                               * identical to the reference's own slicetype_frame_cost with 1 lookahead thread */
    int max_frames;           /* frame handles that may be alive at once (rc.i_lookahead + bframes + 3)   */
    const uint16_t *cost_mv;  /* optional: host pointer to the CENTRE of h->cost_mv[X264_LOOKAHEAD_QP]
                               * (analyse.c:143-157), valid for +-8*mv_range; NULL: rebuilt with the same
                               * formula                                                                  */
} x264cu_cfg;

/* explicit weighted prediction of list 0 of a P evaluation: x264_weight_t (common/mc.h:235-245).
 * enabled == X264CU_WEIGHT_DEVICE: use the result that x264cu_weights_analyse( fenc, fref0 ) left in device memory
 * (no host round trip between the analysis and the search, slicetype.c:859-867) */
#define X264CU_WEIGHT_DEVICE 2
typedef struct { int enabled, scale, denom, offset; } x264cu_weight;

/* Host destinations of one evaluation: the x264_frame_t lookahead fields (common/frame.h:103-127).
 * NULL members are not copied.  Filled when x264cu_flush returns. */
typedef struct
{
    uint16_t *lowres_costs;   /* lowres_costs[b-p0][p1-b]      u16 x mb_count: min(cost,0x3FFF)+(list<<14) */
    int16_t  *mvs[2];         /* lowres_mvs[l][dist-1]         i16 x 2 x mb_count (qpel)                   */
    int32_t  *mv_costs[2];    /* lowres_mv_costs[l][dist-1]    i32 x mb_count                              */
    int32_t  *row_satds;      /* i_row_satds[b-p0][p1-b]       i32 x mb_height                             */
    int32_t  *stats;          /* [0] i_cost_est[b-p0][p1-b] (B: already x100/(120+bias)), [1] i_cost_est_aq,
                               * [2] i_intra_mbs[b-p0], [3] raw (unscaled) cost                            */
    uint16_t *intra_costs;    /* i_intra_cost                  u16 x mb_count                              */
    int32_t  *intra_row_satds;/* i_row_satds[0][0]             i32 x mb_height                             */
    int32_t  *intra_stats;    /* [0] i_cost_est[0][0], [1] i_cost_est_aq[0][0]                             */
} x264cu_out;

/* One frame-cost evaluation of a batch: slicetype_frame_cost(h,a,frames,p0,p1,b) (slicetype.c:836). */
typedef struct
{
    x264cu_frame *fenc, *fref0, *fref1;   /* frames[b], frames[p0], frames[p1] */
    int d0, d1;                           /* b-p0 and p1-b                     */
    x264cu_weight weight;                 /* used only when d1 == 0            */
    x264cu_out out;
} x264cu_job;

/* ---- library / context: x264_opencl_load_library, x264_opencl_lookahead_init/_delete (opencl.h:796-804) ---- */
int          x264cu_device_count( void );
const char  *x264cu_last_error( void );
x264cu_ctx  *x264cu_open( int device, const x264cu_cfg *cfg );
void         x264cu_close( x264cu_ctx *ctx );
/* [0] mb_width [1] mb_height [2] lowres width [3] lowres height [4] device plane stride [5] SM count */
int          x264cu_info( x264cu_ctx *ctx, int *out6 );

/* ---- frames: lazily created device objects of x264_opencl_lowres_init (slicetype-cl.c:101-160),
 *      released by x264_opencl_frame_delete (opencl.c:620) ---- */
x264cu_frame *x264cu_frame_new( x264cu_ctx *ctx );
void          x264cu_frame_free( x264cu_ctx *ctx, x264cu_frame *f );

/* x264_frame_init_lowres (mc.c:458-482) + x264_opencl_lowres_init (slicetype-cl.c:82-282):
 * upload the luma plane (width x height, `stride` bytes per row) and inv_qscale_factor
 * (u16 x mb_count, NULL: 256 everywhere), build the four lowres phase planes with their 32-pixel
 * border, the intra costs and their row/frame sums, and reset the frame's cost/MV caches.
 * `luma`/`inv_qscale` are HOST pointers; the _device variant takes DEVICE pointers (no PCIe copy). */
int x264cu_frame_upload( x264cu_ctx *ctx, x264cu_frame *f, const uint8_t *luma, int stride, const uint16_t *inv_qscale );
int x264cu_frame_upload_device( x264cu_ctx *ctx, x264cu_frame *f, const uint8_t *d_luma, int stride, const uint16_t *d_inv_qscale );

/* ---- adaptive quantisation + weighted-prediction analysis: the two producers right before the path ----
 * x264_adaptive_quant_frame (ratecontrol.c:304-415) followed by x264_frame_init_lowres, the order of
 * encoder.c:3417-3421: upload luma and the interleaved NV12 chroma plane (width x height/2 bytes), compute on the
 * device the per-macroblock AC energy (luma 16x16 + both chroma 8x8), f_qp_offset / f_qp_offset_aq,
 * i_inv_qscale_factor and the frame's i_pixel_sum / i_pixel_ssd, then everything x264cu_frame_upload does.
 * aq_mode 0 (X264_AQ_NONE: sums only), 1 (X264_AQ_VARIANCE), 2 (X264_AQ_AUTOVARIANCE) and 3 (X264_AQ_AUTOVARIANCE_BIASED)
 * are all bit-exact with the reference, floats included.  Modes 2/3 reproduce the host libm's powf(x, 0.125f)
 * (ratecontrol.c:361) through a rounded triple square root plus an exception table built from that libm on first use. */
typedef struct { int mode; float strength; } x264cu_aq_cfg;                      /* rc.i_aq_mode, rc.f_aq_strength */
typedef struct
{
    float    *qp_offset;      /* f_qp_offset == f_qp_offset_aq        float x mb_count                 */
    uint16_t *inv_qscale;     /* i_inv_qscale_factor                  u16 x mb_count                   */
    uint64_t *pixel_stats;    /* {i_pixel_sum[p], i_pixel_ssd[p]} for p = Y,U,V after the mean removal */
} x264cu_aq_out;              /* host destinations, filled when x264cu_flush returns; NULL members are skipped */
int x264cu_frame_upload_aq( x264cu_ctx *ctx, x264cu_frame *f, const uint8_t *luma, int stride, const uint8_t *chroma_nv12, int cstride,
                            const x264cu_aq_cfg *aq, const float *quant_offsets, const x264cu_aq_out *out );
/* the same with DEVICE pointers (no PCIe copy); results stay on the device */
int x264cu_frame_upload_aq_device( x264cu_ctx *ctx, x264cu_frame *f, const uint8_t *d_luma, int stride, const uint8_t *d_chroma_nv12,
                                   int cstride, const x264cu_aq_cfg *aq, const float *d_quant_offsets );
/* For callers that keep x264_adaptive_quant_frame on the host (the unmodified encoder, encoder.c:3417): the frame's
 * i_pixel_sum[0] / i_pixel_ssd[0] (after the mean removal, ratecontrol.c:406-414), which is all that
 * x264cu_weights_analyse needs besides the planes.  Call it after x264cu_frame_upload. */
int x264cu_frame_set_pixel_stats( x264cu_ctx *ctx, x264cu_frame *f, uint64_t pixel_sum, uint64_t pixel_ssd );
/* x264_weights_analyse( h, fenc, fref, b_lookahead = 1 ) (slicetype.c:284-501): luma weight of `fenc` against `fref`
 * (dist = frame distance, 1..bframes+1) from the frames' pixel statistics, the intra costs and -- when the list-0 field
 * has already been searched -- the motion-compensated reference.  Both frames need their pixel statistics (x264cu_frame_upload_aq, or x264cu_frame_set_pixel_stats).
 * The result stays in device memory for a following evaluation with weight.enabled == X264CU_WEIGHT_DEVICE and is
 * copied to `out` (host, may be NULL) when x264cu_flush returns.  numslices: weight_slice_header_cost (:170-189). */
int x264cu_weights_analyse( x264cu_ctx *ctx, x264cu_frame *fenc, x264cu_frame *fref, int dist, int numslices, x264cu_weight *out );

/* slicetype_frame_cost through the GPU hook (slicetype.c:878-897): enqueue, asynchronously, the
 * searches this evaluation still needs (x264_opencl_motionsearch), the mode decision and sums
 * (x264_opencl_finalize_cost) and the copies into `out`.  Evaluations are cached per frame exactly
 * like i_cost_est / the 0x7FFF MV sentinel (slicetype.c:848-867). */
int x264cu_frame_cost( x264cu_ctx *ctx, const x264cu_job *job );
/* x264_opencl_motionsearch (slicetype-cl.c:349): enqueue only the search of one field
 * (frame `fenc` against `fref`, list 0/1, dist = |b - ref|), unless it is cached.  `w` applies to
 * list 0 only; mvs / mv_costs are optional host destinations (i16x2 / i32 per block). */
int x264cu_motionsearch( x264cu_ctx *ctx, x264cu_frame *fenc, x264cu_frame *fref, int list, int dist,
                         const x264cu_weight *w, int16_t *mvs, int32_t *mv_costs );
/* Several searches as ONE batch of wide launches (what x264_opencl_slicetype_prep does with its loop of
 * x264_opencl_motionsearch calls, slicetype-cl.c:676-704); cached fields are skipped.  No host outputs. */
typedef struct { x264cu_frame *fenc, *fref; int list, dist; x264cu_weight weight; } x264cu_search_job;
int x264cu_search_batch( x264cu_ctx *ctx, const x264cu_search_job *jobs, int n_jobs );
/* One-stream sharding across GPUs (BASELINE.json configs[3]): a searched MV field is the one
 * piece of data an evaluation on another GPU needs (fref1->lowres_mvs[0] for the temporal-direct candidate,
 * slicetype.c:629; lowres_mvs / lowres_mv_costs of list 0 for later B evaluations).  x264cu_field_device returns the
 * DEVICE addresses of a field's MVs (i16 x 2 x mb_count) and MV costs (i32 x mb_count) so that the caller can move
 * them between contexts with NCCL / peer copies; x264cu_field_set_searched marks a field that was filled that way
 * (the 0x7FFF sentinel protocol of slicetype.c:855-867). */
int x264cu_field_device( x264cu_ctx *ctx, x264cu_frame *f, int list, int dist, void **d_mvs, void **d_mv_costs );
int x264cu_field_set_searched( x264cu_ctx *ctx, x264cu_frame *f, int list, int dist );
/* x264_opencl_slicetype_prep (slicetype-cl.c:653-712): many evaluations, semantically in array
 * order, executed as a few wide launches (all searches of a pass together, all decisions together). */
int x264cu_frame_cost_batch( x264cu_ctx *ctx, const x264cu_job *jobs, int n_jobs );
/* x264_opencl_flush (slicetype-cl.c:58-70): wait for the stream, then complete the queued copies. */
int x264cu_flush( x264cu_ctx *ctx );
/* Pipelining: every enqueueing call is a batch with an increasing sequence number; up to 3 further batches may
 * be enqueued before the host has to wait for the oldest one.  x264cu_wait completes (device work + host copies)
 * every batch up to `seq` only -- the tail keeps running -- whereas x264cu_flush completes everything. */
uint64_t x264cu_last_batch( x264cu_ctx *ctx );
/* an empty batch: its sequence number completes when everything enqueued so far (kernels and x264cu_memcpy_* copies) has */
uint64_t x264cu_fence( x264cu_ctx *ctx );
int x264cu_wait( x264cu_ctx *ctx, uint64_t seq );
/* number of kernels launched by this context so far */
int64_t x264cu_launch_count( x264cu_ctx *ctx );
/* device-side time of the kernels between the last two flushes, in milliseconds, per kernel class:
 * [0] lowres [1] intra [2] search [3] finalize; also returns their launch counts in n4 (may be NULL) */
int x264cu_kernel_times( x264cu_ctx *ctx, float *ms4, int *n4 );
/* per-class CUDA-event timing of the kernels (off by default: two event records per launch) */
int x264cu_set_profiling( x264cu_ctx *ctx, int on );
/* copy a device lowres plane (incl. border; pitch = lowres width + 64) back to the host: test hook */
int x264cu_frame_get_lowres( x264cu_ctx *ctx, x264cu_frame *f, int plane, uint8_t *out );

/* ---- lookahead window (host C++, x264-i386pic_b200/host/stream.cpp): what one encoder's lookahead
 *      holds (encoder/lookahead.c:68-129).  Each push uploads frame n and evaluates the whole
 *      increment of the B-frame cost matrix, slicetype_frame_cost(p0, n, b) for
 *      n-(bframes+1) <= p0 < b <= n plus the I cost, in the order slicetype_path_cost issues them
 *      (slicetype.c:1288-1331): per p0 the P evaluation, then the B evaluations. ---- */
typedef struct x264cu_stream x264cu_stream;
x264cu_stream *x264cu_stream_open( x264cu_ctx *ctx, int bframes, int want_arrays );
void     x264cu_stream_close( x264cu_stream *s );
/* returns the number of evaluations enqueued (asynchronously) or -1; on_device: device pointers */
int      x264cu_stream_push( x264cu_stream *s, const uint8_t *luma, int stride, const uint16_t *inv_qscale, int on_device );
/* The same push split over `world` GPUs that all hold the window of ONE encode stream (every rank pushes every frame).
 * phase 0: upload frame n and enqueue this rank's share of the new searches -- list-0 field (n -> p0) where
 * p0 % world == rank, list-1 field (b -> n) where b % world == rank.  The caller then exchanges the list-0 fields of
 * frame n between the ranks (x264cu_stream_frame + x264cu_field_device + x264cu_field_set_searched).
 * phase 1: enqueue this rank's share of the evaluations -- P (p0, n, n) where p0 % world == rank, B (p0, n, b) where
 * b % world == rank, I where n % world == rank.  Results of other ranks' evaluations read as -1 in x264cu_stream_cost. */
int      x264cu_stream_push_shard( x264cu_stream *s, const uint8_t *luma, int stride, const uint16_t *inv_qscale, int on_device,
                                   int rank, int world, int phase );
/* phase 0 of frame n+1 may be issued before phase 1 of frame n, so that the exchange of one step overlaps the searches
 * of the next.  x264cu_stream_frame: the handle of frame number `index` (0 = first frame pushed) while it is in the window. */
x264cu_frame *x264cu_stream_frame( x264cu_stream *s, int64_t index );
int      x264cu_stream_sync( x264cu_stream *s );
/* after sync: stats of slicetype_frame_cost(p0, n, b) of the newest frame n, d0 = b-p0, d1 = n-b */
int      x264cu_stream_cost( x264cu_stream *s, int d0, int d1, int32_t *out4 );
/* want_arrays: where the eagerly copied arrays of evaluation (d0, d1) of the most recent push are (host pointers owned by
 * the window, NULL members = not produced; every new MV field is delivered once, with the first evaluation that uses it) */
int      x264cu_stream_arrays( x264cu_stream *s, int d0, int d1, x264cu_out *out );
/* On-demand arrays of slicetype_frame_cost( p0 = b-d0, p1 = b+d1, b ) for frame number b (0 = first frame pushed) while
 * b-d0 .. b+d1 are in the window: the per-evaluation arrays stay on the device and only what a consumer reads crosses
 * PCIe (the decided path of a mini-GOP needs lowres_costs + MV fields of <= bframes+1 evaluations, slicetype.c:1056,
 * 1078-1083, mvpred.c:553; slicetype-cl.c:254-280 copies them eagerly instead).  Computes the evaluation if it was never
 * asked for.  Any subset of `out` may be set; complete after x264cu_stream_sync. */
int      x264cu_stream_fetch( x264cu_stream *s, int64_t b, int d0, int d1, const x264cu_out *out );
int64_t  x264cu_stream_frames( x264cu_stream *s );
int64_t  x264cu_stream_result_bytes( x264cu_stream *s );

/* ---- MB-tree (the consumer right after the path): macroblock_tree_propagate =
 *      mbtree_propagate_cost + mbtree_propagate_list (common/mc.c:511-598) and macroblock_tree_finish
 *      (encoder/slicetype.c:1029-1089) on the device.  The frame walk of macroblock_tree (:1091-1188) and the slicetype decision
 *      logic (:1288-1743) are NOT part of this library: they stay the reference's own host C and
 *      call these entry points / read the cost matrix. ---- */
int x264cu_mbtree_reset( x264cu_ctx *ctx, x264cu_frame *f );                         /* i_propagate_cost = 0 */
int x264cu_mbtree_swap( x264cu_ctx *ctx, x264cu_frame *a, x264cu_frame *b );         /* XCHG of the two i_propagate_cost arrays */
/* frame b = fb evaluated as (d0 = b-p0, d1 = p1-b) propagates into fp0 / fp1; fps_factor as slicetype.c:1065 */
int x264cu_mbtree_propagate( x264cu_ctx *ctx, x264cu_frame *fb, x264cu_frame *fp0, x264cu_frame *fp1, int d0, int d1,
                             float fps_factor, int referenced );
/* qp_offset_aq / qp_offset: host float[mb_count]; qp_offset is read-modify-write (blocks with intra_cost 0 keep their value).
 * qp_offset_aq == NULL: use the device copy that x264cu_frame_upload_aq computed for this frame */
int x264cu_mbtree_finish( x264cu_ctx *ctx, x264cu_frame *f, int fps_factor, float weightdelta, float strength,
                          const float *qp_offset_aq, float *qp_offset );
int x264cu_mbtree_get_propagate( x264cu_ctx *ctx, x264cu_frame *f, uint16_t *out );   /* i_propagate_cost, host u16[mb_count] */

/* ---- batched pixel primitives (config "checkasm-style"): x264_pixel_function_t sad/satd/sa8d/
 *      hadamard_ac (pixel.h:82-103), x264_dct_function_t sub4x4/8x8/16x16_dct (dct.h:28-33),
 *      x264_quant_function_t quant_4x4/8x8 (quant.h:30-33), frame_init_lowres_core (mc.h:303).
 *      Pointers are DEVICE pointers when on_device != 0, else HOST pointers (copied through). ---- */

/* n independent block pairs at byte offsets off1[i] / off2[i] of two planes (checkasm.c:393-397) */
int x264cu_pixel_cmp( x264cu_ctx *ctx, int metric, int size, const uint8_t *pix1, int64_t len1, int stride1,
                      const uint8_t *pix2, int64_t len2, int stride2, int64_t n, const int64_t *off1,
                      const int64_t *off2, int32_t *out, int on_device );
/* every non-overlapping block of pix1 (width x height) against pix2 displaced by (dx,dy) >= 0;
 * pix2 must cover (width+dx) x (height+dy).  out: (height/bh) x (width/bw) scores, raster order. */
int x264cu_pixel_cmp_grid( x264cu_ctx *ctx, int metric, int size, const uint8_t *pix1, int stride1,
                           const uint8_t *pix2, int stride2, int width, int height, int dx, int dy,
                           int32_t *out, int on_device );
/* sad_x3 / sad_x4 (pixel.c:441-464): fenc has stride 16, the refs share `stride` */
int x264cu_sad_x( x264cu_ctx *ctx, int size, int nref, const uint8_t *fenc, const uint8_t *refs, int64_t reflen,
                  int stride, const int64_t *ref_off, int32_t *scores );
int x264cu_hadamard_ac( x264cu_ctx *ctx, int size, const uint8_t *pix, int64_t len, int stride, int64_t n,
                        const int64_t *off, uint64_t *out, int on_device );
int x264cu_hadamard_ac_grid( x264cu_ctx *ctx, int size, const uint8_t *pix, int stride, int width, int height,
                             uint64_t *out, int on_device );
/* bsize 4, 8 or 16; out: n x bsize*bsize coefficients in the reference's order (dct.c:157-205) */
int x264cu_sub_dct( x264cu_ctx *ctx, int bsize, const uint8_t *pix1, int64_t len1, int stride1,
                    const uint8_t *pix2, int64_t len2, int stride2, int64_t n, const int64_t *off1,
                    const int64_t *off2, int16_t *out, int on_device );
int x264cu_sub_dct_grid( x264cu_ctx *ctx, int bsize, const uint8_t *pix1, int stride1, const uint8_t *pix2,
                         int stride2, int width, int height, int16_t *out, int on_device );
/* in-place quant of nblocks blocks of 16 (b8x8 = 0) or 64 coefficients; nz[i] = !!nonzero (quant.c:59-73) */
int x264cu_quant( x264cu_ctx *ctx, int b8x8, int16_t *coefs, int64_t nblocks, const uint16_t *mf,
                  const uint16_t *bias, int32_t *nz, int on_device );
/* frame_init_lowres_core (mc.c:484-507): src must have 2*height+1 rows of 2*width+1 valid bytes */
int x264cu_lowres_core( x264cu_ctx *ctx, const uint8_t *src, int src_stride, uint8_t *dst0, uint8_t *dsth,
                        uint8_t *dstv, uint8_t *dstc, int dst_stride, int width, int height, int on_device );

/* ---- the checkasm surfaces of the fetch / prediction code (x264_mc_functions_t mc.h:267-340, pixel.h intra_mbcmp_x3_8x8c,
 *      quant.h:33).  The lookahead kernels carry in-register forms of the same arithmetic; these expose it per function. ---- */
/* mc_luma == get_ref (common/mc.c:198-249; checkasm.c:1234-1307): block i = w x h pixels at byte offset org[i] of the four
 * phase planes buf + plane_off[0..3], displaced by the quarter-pel vector mv[2i], mv[2i+1]; weight == NULL or !enabled:
 * x264_weight_none; denom == 0 is the opscale_noden form (mc.c:118).  out: n blocks of w*h bytes, packed. */
int x264cu_mc_luma( x264cu_ctx *ctx, const uint8_t *buf, int64_t len, const int64_t plane_off[4], int stride, int64_t n,
                    const int64_t *org, const int16_t *mv, int w, int h, const x264cu_weight *weight, uint8_t *out, int on_device );
/* x264_mc_functions_t.avg[size] (mc.c:49-103; checkasm.c:1325-1353): size = PIXEL_16x16 .. PIXEL_2x2 (0..11), any weight
 * (-63..127: the clip of mc.c:86 acts); 32 = plain rounding average.  out: n blocks of w*h bytes, packed. */
int x264cu_pixel_avg( x264cu_ctx *ctx, int size, const uint8_t *pix1, int64_t len1, int stride1, const uint8_t *pix2, int64_t len2, int stride2,
                      int64_t n, const int64_t *off1, const int64_t *off2, int weight, uint8_t *out, int on_device );
/* mc_weight (mc.c:117-137; checkasm.c:1355-1429) over a width x height plane: the form x264_weight_scale_plane uses
 * (frame.c:825-840); enabled is ignored, denom >= 0 */
int x264cu_mc_weight( x264cu_ctx *ctx, const uint8_t *src, int src_stride, uint8_t *dst, int dst_stride, int width, int height,
                      const x264cu_weight *weight, int on_device );
/* intra_sad_x3_8x8c / intra_satd_x3_8x8c (pixel.c:542-556; checkasm.c:754): res[3i..3i+2] = { DC, H, V } of the 8x8 block at
 * fenc + off1[i]; the predictors read the row above and the column left of fdec + off2[i] (predict.c:221-277) */
int x264cu_intra_x3_8x8c( x264cu_ctx *ctx, int b_satd, const uint8_t *fenc, int64_t len1, int stride1, const uint8_t *fdec, int64_t len2, int stride2,
                          int64_t n, const int64_t *off1, const int64_t *off2, int32_t *res, int on_device );
/* quant_4x4x4 (quant.c:75-86): ngroups x 4 blocks of 16 coefficients in place; nz_mask[g] bit j = block j of group g is non-zero */
int x264cu_quant_4x4x4( x264cu_ctx *ctx, int16_t *coefs, int64_t ngroups, const uint16_t *mf, const uint16_t *bias, int32_t *nz_mask, int on_device );

/* device memory helpers for callers that keep data resident (bench, tests) */
void *x264cu_malloc( x264cu_ctx *ctx, size_t bytes );
void  x264cu_free( x264cu_ctx *ctx, void *p );
/* ---- peer exchange: one encode stream sharded over the GPUs of one box ----
 * The one exchange step of the path: a B evaluation needs fref1->lowres_mvs[0][p1-p0-1] (slicetype.c:629-631), which
 * another rank searched.  Every rank owns a mailbox in its HBM (exported with CUDA IPC); x264cu_peer_push stores this
 * rank's fields straight into all peers' mailboxes over NVLink from one kernel and publishes the step with a
 * release-store of a flag; x264cu_peer_import waits for the flag ON THE DEVICE and copies mailbox -> frame fields.
 * Neither call synchronises with the host.  All ranks must issue the same sequence of push / import calls.
 * handle_out / offset_out of every rank are exchanged by the caller (any host channel) and passed to connect. */
typedef struct x264cu_peer x264cu_peer;
x264cu_peer *x264cu_peer_create( x264cu_ctx *ctx, int rank, int world, int items_per_rank, uint8_t handle_out[64], uint64_t *offset_out );
int  x264cu_peer_connect( x264cu_peer *p, const uint8_t *handles /* world x 64 */, const uint64_t *offsets /* world */ );
/* the same for ranks that live in one process (their contexts may share a device): all[r] = rank r's peer object */
int  x264cu_peer_connect_local( x264cu_peer *p, x264cu_peer *const *all, int world );
int  x264cu_peer_push( x264cu_peer *p, const void *const *src, const int *slot, int n );
int  x264cu_peer_import( x264cu_peer *p, void *const *dst, const int *from, const int *slot, int n );
int  x264cu_peer_error( x264cu_peer *p );
void x264cu_peer_destroy( x264cu_peer *p );

/* page-locked host memory (the reference's CL_MEM_ALLOC_HOST_PTR buffer, opencl.c:380-392): uploads
 * from it and results into it skip the internal staging copy */
void *x264cu_host_alloc( x264cu_ctx *ctx, size_t bytes );
void  x264cu_host_free( x264cu_ctx *ctx, void *p );
int   x264cu_memcpy_h2d( x264cu_ctx *ctx, void *dst, const void *src, size_t bytes );
int   x264cu_memcpy_d2h( x264cu_ctx *ctx, void *dst, const void *src, size_t bytes );
int   x264cu_memcpy_d2d( x264cu_ctx *ctx, void *dst, const void *src, size_t bytes );      /* on the context's compute stream */
/* bracket device work with CUDA events on the context's stream: returns elapsed ms of [begin,end) */
int   x264cu_timer_begin( x264cu_ctx *ctx );
float x264cu_timer_end( x264cu_ctx *ctx );

#ifdef __cplusplus
}
#endif
#endif
