// lookahead.cuh -- device-side descriptors shared between the lookahead kernels and the C ABI.
#pragma once
#include "common.cuh"

// CTA tiles of the thread-per-block kernels in blocks; shared by the kernels (la_block.cuh) and the host-side warp-tile
// table (capi.cu build_warp_tiles).  A CTA is TILE_W * TILE_H threads; a warp covers TILE_W x (32 / TILE_W) blocks.
#ifndef X264CU_S_TILE_W
#define X264CU_S_TILE_W 32
#define X264CU_S_TILE_H 8
#endif
// the two compile-time plane pitches of the search / decision kernels (capi.cu picks the smaller one that fits)
#ifndef X264CU_PITCH_A
#define X264CU_PITCH_A 1024
#define X264CU_PITCH_B 2048
#endif
#ifndef X264CU_F_TILE_W
#define X264CU_F_TILE_W 16
#define X264CU_F_TILE_H 8
#endif

// Geometry and policy of one context; passed by value to every kernel.
struct LaParams
{
    int mbw, mbh, nmb;          // 8x8 lowres blocks (== 16x16 macroblocks of the full frame)
    int lw, lh;                 // lowres plane size
    int stride;                 // device plane pitch (bytes); planes carry a 32-px border
    int64_t plane_size;         // bytes between the four phase planes F,H,V,C of one frame
    int mv_range2;              // 2 * analyse.i_mv_range (slicetype.c:550)
    int me_range;
    int me_method;              // 0 dia, 1 hex (after lowres_context_init, slicetype.c:50-59)
    int b_satd;                 // param subme > 1: mbcmp == SATD, la subme 4; else SAD, la subme 2
    int do_edges;
    int aq;
    int weighted_bipred;
    int lanes_per_block;        // tuning: 0 auto, 1 thread-per-block kernels, 8 eight-lanes-per-block kernels
    int cost_mv_n;              // table covers |d| <= cost_mv_n
    const uint16_t *cost_mv;    // device table indexed by |mvd| in qpel
    // warp tiles of the thread-per-block kernels (capi.cu build_warp_tiles): entry = x0 | y0 << 12 | log2(width) << 24;
    // a warp covers width x (32/width) blocks.  Frame widths that are not a multiple of the tile width are packed with
    // narrower, taller tiles instead of leaving lanes idle (1080p: 120 block columns = 3 x 32 + 16 + 8)
    const uint32_t *tiles_s; int ntiles_s;      // search kernels: CTA = 4 consecutive tiles
    const uint32_t *tiles_f; int ntiles_f;      // decision kernels
};

// One motion-search pass over one (fenc, ref, list) field: x264_me_search for every block.
struct SearchJob
{
    const uint8_t *fenc;        // F plane of the frame being predicted, pixel (0,0)
    const uint8_t *ref;         // F plane of the reference, pixel (0,0); H,V,C follow at plane_size
    const uint8_t *ref_w;       // fullpel plane: `ref`, or its weighted copy (slicetype.c:612-614)
    const int16_t *mv_in;       // MV field of the previous pass (NULL: the zero field)
    int16_t *mv_out;            // MV field written by this pass (x,y int16 per block, qpel)
    int32_t *cost_out;          // lowres_mv_costs written by this pass
    int wscale, wdenom, woffset;// wdenom < 0: unweighted
    const int *wdev;            // != NULL: the weight is the device-side result of x264cu_weights_analyse
                                // ([0] enabled [1] scale [2] denom [3] offset); ref_w is used only when enabled
};

// Mode decision + sums of one (b,p0,p1) evaluation: the tail of slicetype_mb_cost.
struct FinalJob
{
    const uint8_t *fenc, *ref0, *ref1;      // F planes, pixel (0,0)
    const int16_t *mv[2];                   // searched fields (NULL when the list is unused)
    const int32_t *mvcost[2];
    const int16_t *dmv_src;                 // fref1->lowres_mvs[0][p1-p0-1] if already searched, else NULL
    const uint16_t *intra;                  // fenc intra costs
    const uint16_t *inv_qscale;
    uint16_t *lowres_costs;                 // out, per block
    int32_t *row_satds;                     // out, mbh sums (pre-zeroed)
    int32_t *stats;                         // out: cost_est, cost_est_aq, intra_mbs (pre-zeroed)
    int b_bidir, b_intra_only;              // b < p1 ; p0 == p1
    int dist_scale_factor, bipred_weight;
};

// One contiguous copy of the result gather (device -> device staging block).
struct GatherSeg { const uint8_t *src; unsigned long long dst_off; unsigned bytes; };
void launch_gather( const GatherSeg *d_segs, int nsegs, uint8_t *dst_base, cudaStream_t st );

void launch_lowres( const uint8_t *luma, int stride, int width, int height, uint8_t *planes, const LaParams &P,
                    int32_t *stats, int nstats, uint16_t *inv_qscale, unsigned long long *pixstats, cudaStream_t st );
void launch_weight_plane( const uint8_t *src, uint8_t *dst, int64_t bytes, int scale, int denom, int offset, cudaStream_t st );
void launch_intra( const uint8_t *planeF, const uint16_t *inv_qscale, uint16_t *intra, int32_t *row_satds,
                   int32_t *stats, const LaParams &P, cudaStream_t st );
void launch_search( const SearchJob *d_jobs, int njobs, bool any_weighted, bool first_pass, const LaParams &P, cudaStream_t st );
// d_chunks: (first job, count) per CTA row of the thread-per-block kernel; every chunk is homogeneous (all bidir with the
// same fenc / ref1 / list-1 field, or all non-bidir), see capi.cu group_finals
void launch_final( const FinalJob *d_jobs, int njobs, const int2 *d_chunks, int nchunks, const LaParams &P, cudaStream_t st );
#ifndef X264CU_F_CHUNK
#define X264CU_F_CHUNK 8
#endif
// returns the number of launches; pow_x / pow_v / pow_n: the powf exception table of aq-mode 2 / 3 (aq_weights.cu), red3: 3 floats of scratch
int  launch_aq( const uint8_t *luma, int ls, const uint8_t *chroma, int cs, int W, int H, const LaParams &P, int mode, float strength,
                const float *quant_offsets, const float *log2_lut, const uint8_t *exp2_lut, float *qp_offset, uint16_t *inv_qscale,
                unsigned long long *pixstats, const u32 *pow_x, const float *pow_v, int pow_n, float *red3, cudaStream_t st );
void launch_weights_analyse( const uint8_t *fencF, const uint8_t *refF, const int16_t *mvs, const uint16_t *intra,
                             const unsigned long long *st_fenc, const unsigned long long *st_ref, int *wres, int numslices,
                             const LaParams &P, cudaStream_t st );
void launch_weight_plane_dev( const uint8_t *src, uint8_t *dst, int64_t bytes, const int *wres, cudaStream_t st );
int  launch_search_serial( const SearchJob *d_jobs, int njobs, const LaParams &P, int *d_err, int max_resident_ctas, cudaStream_t st );
int  search_serial_max_ctas_per_sm();
void launch_search_waves( const SearchJob *d_jobs, int njobs, const LaParams &P, cudaStream_t st );

// One macroblock_tree_propagate step (slicetype.c:1051-1089) of frame b towards its references.
struct MbtreeJob
{
    const uint16_t *intra, *lowres_costs, *inv_qscale;   // of frame b: i_intra_cost, lowres_costs[b-p0][p1-b], i_inv_qscale_factor
    const u32 *prop_b;                                     // frame b's own accumulated propagate cost (read when referenced)
    u32 *prop_ref[2];                                      // accumulators of frames p0 / p1
    const int16_t *mvs[2];                                 // lowres_mvs[l][dist-1] of frame b (NULL: list unused)
    int bipred_weight[2];
    float fps_factor;
    int referenced;
};
void launch_mbtree_propagate( const MbtreeJob &J, const LaParams &P, cudaStream_t st );
void launch_mbtree_finish( const uint16_t *intra, const uint16_t *inv_qscale, const u32 *prop, const float *qp_offset_aq, float *qp_offset,
                           const float *log2_lut, int fps_factor, float weightdelta, float strength, int nmb, cudaStream_t st );
