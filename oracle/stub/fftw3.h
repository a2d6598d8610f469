/* TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * Minimal stand-in for <fftw3.h> so that the reference's hot-path sources,
 * which include it only for the `fftw_complex` type (gen-pk.h:21,
 * powerspectrum.c:15), compile in an image where FFTW3 is not installed.
 * It declares the type and the plan prototypes the reference names
 * (gen-pk.cpp:178-193,233,363-364); none of the functions is defined or
 * called by the objects that go into oracle/_ref/libgenpk_ref.so.
 */
#ifndef GENPK_ORACLE_FFTW3_STUB_H
#define GENPK_ORACLE_FFTW3_STUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double fftw_complex[2];
typedef struct fftw_plan_s *fftw_plan;
#define FFTW_ESTIMATE (1U << 6)
void *fftw_malloc(size_t n);
void fftw_free(void *p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in, fftw_complex *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
#ifdef __cplusplus
}
#endif
#endif
