/* Hand-written stand-in for the reference's generated x264_config.h (oracle build only). */
#define X264_GPL           1
#define X264_INTERLACED    1
#define X264_BIT_DEPTH     8
#define X264_CHROMA_FORMAT 0
#define X264_VERSION ""
#define X264_POINTVER "0.164.x"
