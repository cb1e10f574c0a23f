/* oracle/x264_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the x264 lookahead cost path that the CUDA product implements.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * Parity status: PINNED -- every function here is checked against the unmodified reference
 * (oracle/_ref/libx264ref.so, built from /root/reference by oracle/Makefile) by
 * tests/test_oracle_vs_ref.py, and against the golden vectors under tests/golden/ that were
 * generated from that reference build (tests/golden/make_golden.py).
 */
#ifndef X264_ORACLE_H
#define X264_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* block sizes in the order of the reference's PIXEL_* enum (common/pixel.h:41-52) */
enum { XO_16x16, XO_16x8, XO_8x16, XO_8x8, XO_8x4, XO_4x8, XO_4x4, XO_4x16 };

/* ---- primitives (pixel.c / dct.c / quant.c / mc.c / predict.c) ---- */
int      xo_sad( int size, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb );
int      xo_satd( int size, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb );
int      xo_sa8d( int size, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb );
uint64_t xo_hadamard_ac( int size, const uint8_t *p, intptr_t stride );
int      xo_pixel_cmp_batch( int kind, int size, const uint8_t *base1, intptr_t s1, const uint8_t *base2, intptr_t s2,
                             int64_t n, const int64_t *off1, const int64_t *off2, int32_t *out );
int      xo_hadamard_ac_batch( int size, const uint8_t *base, intptr_t stride, int64_t n, const int64_t *off, uint64_t *out );
void     xo_sub_dct( int bsize, int16_t *out, const uint8_t *a, intptr_t sa, const uint8_t *b, intptr_t sb );
int      xo_sub_dct_batch( int bsize, const uint8_t *base1, intptr_t s1, const uint8_t *base2, intptr_t s2,
                           int64_t n, const int64_t *off1, const int64_t *off2, int16_t *out );
int      xo_quant( int16_t *coef, int n, const uint16_t *mf, const uint16_t *bias );
int      xo_quant_batch( int b8x8, int16_t *coefs, int64_t nblocks, const uint16_t *mf, const uint16_t *bias, int32_t *nz );
void     xo_lowres_core( const uint8_t *src, intptr_t sstride, uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc,
                         intptr_t dstride, int width, int height );
void     xo_mc_luma( uint8_t *dst, intptr_t ds, const uint8_t *const planes[4], intptr_t stride,
                     int mvx, int mvy, int w, int h, int scale, int denom, int offset );
void     xo_avg( uint8_t *dst, intptr_t ds, const uint8_t *a, intptr_t as, const uint8_t *b, intptr_t bs,
                 int w, int h, int weight );
void     xo_mc_weight( uint8_t *dst, intptr_t ds, const uint8_t *src, intptr_t ss, int width, int height, int scale, int denom, int offset );
void     xo_intra_x3_8x8c( int b_satd, const uint8_t *fenc, intptr_t sf, const uint8_t *fdec, intptr_t sd, int res[3] );
int      xo_intra_x3_8x8c_batch( int b_satd, const uint8_t *p1, intptr_t s1, const uint8_t *p2, intptr_t s2, int64_t n,
                                 const int64_t *off1, const int64_t *off2, int32_t *res );
/* intra cost of one 8x8 lowres block given its padded plane (slicetype.c:714-747, before AQ) */
int      xo_intra_cost_8x8( const uint8_t *src, intptr_t stride, int b_satd_flavour );
void     xo_cost_mv_table( uint16_t *centre, int n );   /* analyse.c:153-157,186-188 with lambda 1 */

/* ---- the lookahead ---- */
typedef struct
{
    int width, height;      /* luma size before mod-16 padding */
    int bframes;            /* x264_param_t.i_bframe */
    int subme;              /* x264_param_t.analyse.i_subpel_refine (>1: SATD/HEX/subme-4 flavour) */
    int me_method;          /* x264_param_t.analyse.i_me_method (0 dia, 1 hex, ...) */
    int me_range;           /* 16 */
    int mv_range;           /* x264_param_t.analyse.i_mv_range after validation (level-derived) */
    int do_edges;           /* b_mb_tree || vbv (slicetype.c:823) */
    int aq;                 /* rc.i_aq_mode != 0 */
    int weighted_bipred;    /* analyse.b_weighted_bipred */
    int bframe_bias;        /* i_bframe_bias */
    int search_passes;      /* K: Jacobi passes of the dependency-free schedule (>=1) */
} xo_cfg;

typedef struct xo_ctx xo_ctx;

xo_ctx *xo_open( const xo_cfg *cfg );
void    xo_close( xo_ctx * );
int     xo_info( xo_ctx *, int *out /*[4]: mb_w, mb_h, lowres_w, lowres_h*/ );
int     xo_put_frame( xo_ctx *, int idx, const uint8_t *luma, int stride, const uint16_t *inv_qscale );
int     xo_get_lowres( xo_ctx *, int idx, int plane, uint8_t *out );   /* incl. 32-px border, pitch w+64 */
/* weight: denom < 0 -> none.  Applies to list 0 of a P evaluation (b == p1) only. */
int     xo_frame_cost( xo_ctx *, int p0, int p1, int b, int w_scale, int w_denom, int w_offset );
int     xo_get( xo_ctx *, int idx, int what, int i, int j, void *out );   /* same codes as ref_la_get */
/* counters of the last xo_frame_cost: [0] fullpel SAD candidates, [1] subpel SAD, [2] SATD 8x8 */
void    xo_counters( xo_ctx *, int64_t *out );

/* ---- MB-tree: mc.c:511-598, slicetype.c:1029-1089 ---- */
void    xo_mbtree_propagate_cost( int16_t *dst, const uint16_t *propagate_in, const uint16_t *intra_costs, const uint16_t *inter_costs,
                                  const uint16_t *inv_qscales, float fps, int len );
void    xo_mbtree_propagate_list( int mbw, int mbh, uint16_t *ref_costs, const int16_t (*mvs)[2], const int16_t *propagate_amount,
                                  const uint16_t *lowres_costs, int bipred_weight, int mb_y, int len, int list );
int     xo_mbtree_reset( xo_ctx *, int idx );
int     xo_mbtree_swap( xo_ctx *, int a, int b );
int     xo_mbtree_get( xo_ctx *, int idx, uint16_t *out );
int     xo_mbtree_propagate( xo_ctx *, int p0, int p1, int b, float fps_factor, int referenced );
int     xo_mbtree_finish( xo_ctx *, int idx, int fps_factor, float weightdelta, float strength, const float *qp_offset_aq, float *qp_offset );

/* ---- adaptive quantisation + weighted-prediction analysis: ratecontrol.c:226-415 (aq-mode 0/1), slicetype.c:62-501 ---- */
/* x264_adaptive_quant_frame followed by x264_frame_init_lowres (encoder.c:3417-3421); chroma is the interleaved NV12 plane */
int     xo_put_frame_aq( xo_ctx *, int idx, const uint8_t *luma, int stride, const uint8_t *chroma_nv12, int cstride,
                         int aq_mode, float aq_strength, const float *quant_offsets );
/* f_qp_offset (float x mb_count), i_inv_qscale_factor (u16 x mb_count), {i_pixel_sum[p], i_pixel_ssd[p]} p = 0..2 */
int     xo_get_aq( xo_ctx *, int idx, float *qp_offset, uint16_t *inv_qscale, uint64_t *sums6 );
/* x264_weights_analyse( h, frames[b], frames[ref], 1 ); out: enabled, scale, denom, offset, minscore, origscore */
int     xo_weights_analyse( xo_ctx *, int b, int ref, int numslices, int *out6 );
void    xo_exp2_lut( uint8_t *out64 );

#ifdef __cplusplus
}
#endif
#endif
