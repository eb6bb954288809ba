/* mcgvc.h -- C ABI of the B200-native MaskCycleGAN-VC conv engine (libmcgvc.so).
 * Work in progress header: kernel-level debug entry points first; network-level entry points follow. */
#ifndef MCGVC_H
#define MCGVC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* mcgvc_last_error(void);

int mcgvc_debug_conv(const void* a_hi, const void* a_lo, int aC, int aX, int aY, int aP, int aB,
                     const void* w_hi, const void* w_lo, int wK, int wN, int wT, int oX, int oY,
                     int oB, int nTaps, const int8_t* taps4, float* out, long long sB,
                     long long sY, long long sX, int nSplit, long long sNhi, const float* bias,
                     const float* addsrc, int nPass, int backend, int blockN, void* stream);

int mcgvc_debug_wgrad(const void* z_hi, const void* z_lo, int zC, int zX, int zY, int zB,
                      const void* x_hi, const void* x_lo, int xC, int xX, int xY, int xP, int xB,
                      int pX, int pY, int pB, int nTaps, const int8_t* taps4,
                      const int8_t* ztaps4, float* dw, int cTile, int splitK, int nPass,
                      int backend, void* stream);

#ifdef __cplusplus
}
#endif
#endif
