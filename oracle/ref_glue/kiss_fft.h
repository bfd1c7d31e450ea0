/* TEST INFRASTRUCTURE ONLY: declaration-only stand-in for the un-vendored
 * kissfft header (deps/bootstrap.json pins mborgerding/kissfft@7bce4153); only
 * blue_noise2d_tex() uses it and that function is out of scope. */
#ifndef ORACLE_KISS_FFT_H
#define ORACLE_KISS_FFT_H
#include <stddef.h>
typedef struct { float r, i; } kiss_fft_cpx;
typedef struct kiss_fft_state *kiss_fft_cfg;
kiss_fft_cfg kiss_fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);
void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout);
#define kiss_fft_free free
#endif
