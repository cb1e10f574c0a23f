// primitives.cuh -- launchers of the batched pixel / dct / quant primitives (primitives.cu).
#pragma once
#include "common.cuh"

void launch_fill_u16( uint16_t *dst, int value, int n, cudaStream_t st );

// metric: 0 sad, 1 satd, 2 sa8d; size: PIXEL_* order
int launch_pixel_cmp( int metric, int size, const uint8_t *p1, int s1, const uint8_t *p2, int s2,
                      int64_t n, const int64_t *off1, const int64_t *off2, int32_t *out, cudaStream_t st );
int launch_pixel_cmp_grid( int metric, int size, const uint8_t *p1, int s1, const uint8_t *p2, int s2,
                           int width, int height, int32_t *out, cudaStream_t st );
int launch_hadamard_ac( int size, const uint8_t *p, int stride, int64_t n, const int64_t *off, uint64_t *out, cudaStream_t st );
int launch_hadamard_ac_grid( int size, const uint8_t *p, int stride, int width, int height, uint64_t *out, cudaStream_t st );
int launch_sub_dct( int bsize, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n,
                    const int64_t *off1, const int64_t *off2, int16_t *out, cudaStream_t st );
int launch_sub_dct_grid( int bsize, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height,
                         int16_t *out, cudaStream_t st );
int launch_quant( int b8x8, int16_t *coefs, int64_t nblocks, const uint16_t *mf, const uint16_t *bias, int32_t *nz, cudaStream_t st );
int launch_lowres_core( const uint8_t *src, int sstride, uint8_t *d0, uint8_t *dh, uint8_t *dv, uint8_t *dc,
                        int dstride, int width, int height, cudaStream_t st );

// primitives_mc.cu: the checkasm surfaces of the fetch / prediction code (mc_luma == get_ref, avg, weight, intra x3, quant_4x4x4)
int launch_mc_luma( const uint8_t *buf, const int64_t plane_off[4], int stride, int64_t n, const int64_t *org, const int16_t *mv,
                    int w, int h, int wscale, int wdenom, int woffset, uint8_t *out, cudaStream_t st );
int launch_pixel_avg( const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n, const int64_t *off1, const int64_t *off2,
                      int w, int h, int weight, uint8_t *out, cudaStream_t st );
int launch_mc_weight( const uint8_t *src, int sstride, uint8_t *dst, int dstride, int width, int height, int scale, int denom, int offset, cudaStream_t st );
int launch_intra_x3_8x8c( int b_satd, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int64_t n, const int64_t *off1, const int64_t *off2,
                          int32_t *res, cudaStream_t st );
int launch_quant_4x4x4( int16_t *coefs, int64_t ngroups, const uint16_t *mf, const uint16_t *bias, int32_t *nza, cudaStream_t st );
