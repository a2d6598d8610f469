/* TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" doorways into the UNMODIFIED reference objects that
 * oracle/Makefile compiles from /root/reference into oracle/_ref/.  The
 * reference declares fieldize(), read_fieldize(), nexttwo(), type_str() and
 * print_pk() with C++ linkage (gen-pk.h:35-93), so ctypes cannot reach them
 * without these one-line trampolines.  No arithmetic lives here.
 */
#include "gen-pk.h"
#include <string.h>
#include <string>

using namespace GadgetReader;

extern "C" {

/* -> fieldize(), fieldize.cpp:46 */
int ref_fieldize(double boxsize, int dims, double *out, int64_t n, float *positions, float *masses, double mass, int extra)
{
    return fieldize(boxsize, dims, out, n, positions, masses, mass, extra);
}

/* -> nexttwo(), utils.cpp:35 */
int ref_nexttwo(int n) { return nexttwo(n); }

/* -> type_str(), utils.cpp:57 */
void ref_type_str(int type, char *out3)
{
    std::string s = type_str(type);
    strncpy(out3, s.c_str(), 3);
    out3[2] = 0;
}

/* -> print_pk(), utils.cpp:7 */
int ref_print_pk(const char *filename, int nrbins, double *keffs, double *power, int *count)
{
    return print_pk(std::string(filename), nrbins, keffs, power, count);
}

/* GSnap handle helpers (gadgetreader.hpp:190-293) */
void *ref_snap_open(const char *base) { return new GSnap(std::string(base), false); }
void ref_snap_close(void *s) { delete (GSnap *)s; }
int ref_snap_numfiles(void *s) { return ((GSnap *)s)->GetNumFiles(); }
int64_t ref_snap_npart(void *s, int type) { return ((GSnap *)s)->GetNpart(type); }
double ref_snap_mass(void *s, int type) { return ((GSnap *)s)->GetHeader().mass[type]; }
double ref_snap_box(void *s) { return ((GSnap *)s)->GetHeader().BoxSize; }
double ref_snap_redshift(void *s) { return ((GSnap *)s)->GetHeader().redshift; }
double ref_snap_omega0(void *s) { return ((GSnap *)s)->GetHeader().Omega0; }
/* -> GSnap::GetBlock(), gadgetreader.cpp:471 */
int64_t ref_snap_getblock(void *s, const char *name, void *block, int64_t n, int64_t start, int skip_type)
{
    return ((GSnap *)s)->GetBlock(std::string(name), block, n, start, skip_type);
}

/* -> read_fieldize(), read_fieldize.cpp:18 */
int ref_read_fieldize(double *field, void *s, int type, double box, int dims, double *total_mass)
{
    return read_fieldize(field, (GSnap *)s, type, box, dims, total_mass);
}

} /* extern "C" */
