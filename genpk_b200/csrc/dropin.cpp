// libgenpk_dropin.so -- the reference's own link-time symbols for the P(k) hot path,
// implemented on the GPU through the C ABI of libgenpk_cuda.so.
//
// gen-pk links `gen-pk.o read_fieldize.o utils.o read_fieldize_bigfile.o` against
// fieldize.o, powerspectrum.o and -lfftw3 (Makefile:17,44,49-50).  This library
// exports exactly what those three provide to the rest of the program:
//
//   int    fieldize(double,int,double*,int64_t,float*,float*,double,int)   C++ linkage, gen-pk.h:93
//   double invwindow(int64_t,int64_t,int64_t,int64_t)                      extern "C",  gen-pk.h:104
//   int    powerspectrum(int64_t,fftw_complex*,fftw_complex*,int,double*,int*,double*,double,double)   gen-pk.h:119
//   fftw_malloc fftw_free fftw_init_threads fftw_plan_with_nthreads
//   fftw_plan_dft_r2c_3d fftw_execute fftw_destroy_plan                    gen-pk.cpp:176-193,233,363-364
//
// so relinking gen-pk with `-lgenpk_dropin -lgenpk_cuda` instead of those objects and
// -lfftw3 moves the path to the GPU without touching a source line.
//
// The grid stays in HBM.  fftw_plan_dft_r2c_3d(d,d,d,field,field,..) -- the in-place
// cubic plan gen-pk makes once per field (gen-pk.cpp:193,270,321) -- registers the host
// pointer `field` and creates a device grid for it.  fieldize(out == field, extra == 1)
// then deposits straight into that device grid (particles go up, the 8*d^3-byte grid
// never crosses PCIe), fftw_execute only notes that the transform is wanted, and
// powerspectrum(field, ...) runs transform and binning as one fused call
// (genpk_fft_power / genpk_fft_power_cross: the x-transformed spectrum is never written)
// and hands back the three small arrays.  After powerspectrum has consumed a field its device grid is cleared, which is
// what the caller's memset(field, 0, ...) before the next particle type (gen-pk.cpp:208)
// means for the device copy.  The host bytes behind a registered pointer are never
// read or written: gen-pk itself never looks at them.  A caller that does (test.cpp
// reads field[] after fieldize) either uses unregistered buffers -- every entry point
// then falls back to the host-buffer shims of genpk_cuda.h, which round-trip the grid
// -- or sets GENPK_DROPIN_MIRROR=1, which copies the device grid back to the host
// pointer after every fieldize / fftw_execute (exact reference semantics, PCIe-bound).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "../../include/genpk_cuda.h"

typedef double fftw_complex[2];

namespace {

enum State { ZERO, REAL, PENDING, SPECTRUM };   // PENDING: fftw_execute was called, the transform has not run yet

struct Field {
    genpk_ctx *ctx = nullptr;
    int dims = 0;
    State state = ZERO;
    bool device_resident = false;   // in-place cubic plan: the grid lives in HBM
    double *in = nullptr;
    fftw_complex *out = nullptr;
};

std::mutex g_lock;
std::map<const void *, Field *> g_fields;   // keyed by the host pointer given to the plan

bool mirror() { const char *e = getenv("GENPK_DROPIN_MIRROR"); return e && *e && *e != '0'; }
// fftw_execute on a device-resident field only notes the request: gen-pk hands the field to nothing but
// powerspectrum() afterwards (gen-pk.cpp:233-234, 295-297, 345-348), which then runs transform and binning
// as one fused call.  GENPK_DROPIN_EAGER=1 (or the mirror mode) transforms at once.
bool eager() { const char *e = getenv("GENPK_DROPIN_EAGER"); return mirror() || (e && *e && *e != '0'); }

Field *lookup(const void *p)
{
    std::lock_guard<std::mutex> l(g_lock);
    auto it = g_fields.find(p);
    return it == g_fields.end() ? nullptr : it->second;
}

void complain(const char *where) { fprintf(stderr, "genpk drop-in: %s: %s\n", where, genpk_last_error()); }

}  // namespace

// ---- fieldize.cpp:46 (C++ linkage, as in gen-pk.h:93) ---------------------------------
int fieldize(double boxsize, int dims, double *out, int64_t segment_particles, float *positions, float *masses,
             double mass, int extra)
{
    Field *f = lookup(out);
    if (f && f->device_resident && f->dims == dims && extra == 1 && f->state != SPECTRUM && f->state != PENDING) {
        if (genpk_deposit(f->ctx, 0, positions, masses, segment_particles, mass, boxsize, 0)) {
            complain("fieldize");
            return 1;
        }
        f->state = REAL;
        if (mirror() && (genpk_grid_download(f->ctx, 0, out) != 0))
            complain("fieldize (mirror)");
        return 0;
    }
    if (genpk_fieldize(boxsize, dims, out, segment_particles, positions, masses, mass, extra)) {
        complain("fieldize");
        return 1;
    }
    return 0;
}

extern "C" {

// ---- fieldize.cpp:125 -------------------------------------------------------------------
double invwindow(int64_t kx, int64_t ky, int64_t kz, int64_t n) { return genpk_invwindow(kx, ky, kz, n); }

// ---- powerspectrum.c:35 -----------------------------------------------------------------
int powerspectrum(int64_t dims, fftw_complex *outfield, fftw_complex *outfield2, int nrbins, double *power, int *count,
                  double *keffs, double total_mass, double total_mass2)
{
    Field *a = lookup(outfield), *b = outfield2 == outfield ? a : lookup(outfield2);
    if (a && b && a->device_resident && b->device_resident && a->state == PENDING && b->state == PENDING &&
        a->dims == dims && b->dims == dims) {
        // transform(s) + binning in one call on the fused path
        int rc;
        if (a == b)
            rc = genpk_fft_power(a->ctx, 0, nrbins, power, count, keffs, total_mass, total_mass2);
        else
            rc = genpk_fft_power_cross(a->ctx, 0, b->ctx, 0, nrbins, power, count, keffs, total_mass, total_mass2);
        if (rc) {
            complain("powerspectrum");
            return 1;
        }
        genpk_grid_zero(a->ctx, 0);
        a->state = ZERO;
        if (b != a) {
            genpk_grid_zero(b->ctx, 0);
            b->state = ZERO;
        }
        return 0;
    }
    // a deferred transform whose partner is not deferred (or of another size): run it now
    for (Field *f : {a, b})
        if (f && f->device_resident && f->state == PENDING) {
            if (genpk_fft(f->ctx, 0))
                complain("powerspectrum (deferred transform)");
            f->state = SPECTRUM;
        }
    if (a && b && a->device_resident && b->device_resident && a->state == SPECTRUM && b->state == SPECTRUM &&
        a->dims == dims && b->dims == dims) {
        int rc;
        if (a == b)
            rc = genpk_power(a->ctx, 0, 0, nrbins, power, count, keffs, total_mass, total_mass2);
        else
            rc = genpk_power_dev(a->ctx, genpk_grid_device_ptr(a->ctx, 0), genpk_grid_device_ptr(b->ctx, 0), nrbins, power,
                                 count, keffs, total_mass, total_mass2);
        if (rc) {
            complain("powerspectrum");
            return 1;
        }
        // consumed: the caller's next memset(field, 0) (gen-pk.cpp:208) is mirrored on the device
        genpk_grid_zero(a->ctx, 0);
        a->state = ZERO;
        if (b != a) {
            genpk_grid_zero(b->ctx, 0);
            b->state = ZERO;
        }
        return 0;
    }
    if (genpk_powerspectrum(dims, (const double *)outfield, (const double *)outfield2, nrbins, power, count, keffs,
                            total_mass, total_mass2)) {
        complain("powerspectrum");
        return 1;
    }
    return 0;
}

// ---- the FFTW3 calls of gen-pk.cpp ---------------------------------------------------------
void *fftw_malloc(size_t n)
{
    void *p = nullptr;
    return posix_memalign(&p, 64, n ? n : 64) == 0 ? p : nullptr;
}
void fftw_free(void *p) { free(p); }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}

typedef struct fftw_plan_s *fftw_plan;

fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in, fftw_complex *out, unsigned)
{
    Field *f = new Field();
    f->dims = n0;
    f->in = in;
    f->out = out;
    if (n0 == n1 && n1 == n2 && (void *)in == (void *)out && n0 > 0) {
        f->ctx = genpk_create(n0, -1, 0);
        if (!f->ctx || genpk_grid_zero(f->ctx, 0)) {
            complain("fftw_plan_dft_r2c_3d");
            if (f->ctx) genpk_destroy(f->ctx);
            delete f;
            return nullptr;
        }
        f->device_resident = true;
        std::lock_guard<std::mutex> l(g_lock);
        g_fields[in] = f;
    } else if (!(n0 == n1 && n1 == n2)) {
        fprintf(stderr, "genpk drop-in: only the cubic r2c plans of gen-pk.cpp are supported (%d,%d,%d)\n", n0, n1, n2);
        delete f;
        return nullptr;
    }
    return reinterpret_cast<fftw_plan>(f);
}

void fftw_execute(const fftw_plan p)
{
    Field *f = reinterpret_cast<Field *>(p);
    if (!f)
        return;
    if (f->device_resident) {
        // nothing was deposited on the device since the grid was cleared: the caller filled the
        // host buffer itself (test.cpp:64-75), so that is what gets transformed
        if (f->state == ZERO && genpk_grid_upload(f->ctx, 0, f->in))
            complain("fftw_execute (upload)");
        if (!eager()) {
            f->state = PENDING;
            return;
        }
        if (genpk_fft(f->ctx, 0))
            complain("fftw_execute");
        f->state = SPECTRUM;
        if (mirror() && genpk_grid_download(f->ctx, 0, (double *)f->out))
            complain("fftw_execute (mirror)");
        return;
    }
    // out-of-place cubic plan: host round trip through the in-place device transform
    const size_t d = (size_t)f->dims, fd = 2 * (d / 2 + 1);
    double *buf = (double *)f->out;
    for (size_t r = 0; r < d * d; r++)
        memmove(buf + r * fd, f->in + r * d, d * sizeof(double));
    if (genpk_r2c_3d(f->dims, buf))
        complain("fftw_execute");
}

void fftw_destroy_plan(fftw_plan p)
{
    Field *f = reinterpret_cast<Field *>(p);
    if (!f)
        return;
    if (f->device_resident) {
        std::lock_guard<std::mutex> l(g_lock);
        g_fields.erase(f->in);
    }
    if (f->ctx)
        genpk_destroy(f->ctx);
    delete f;
}

}  // extern "C"
