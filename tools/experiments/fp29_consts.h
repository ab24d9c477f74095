// GENERATED (tools inline): BLS12-381 Fr in 29-bit limbs
#define AVRF_P29_BAND {{0x1u, 0x1ffffff8u, 0x1f96ffbfu, 0x1b4805ffu, 0x1d80553bu, 0xc0404d0u, 0x1520cce7u, 0xa6533afu, 0x73eda7u}, 0x1fffffffu}
