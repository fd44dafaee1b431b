/* TEST INFRASTRUCTURE ONLY (oracle). The reference's utilities.cpp links against
 * its libpng wrapper (io_png.h:25-32); libpng headers are not in this image and
 * the oracle never touches files, so these entry points just fail. */
#include <stddef.h>
float *read_png_f32(const char *fname, size_t *nx, size_t *ny, size_t *nc)
{ (void) fname; (void) nx; (void) ny; (void) nc; return NULL; }
int write_png_f32(const char *fname, const float *data, size_t nx, size_t ny, size_t nc)
{ (void) fname; (void) data; (void) nx; (void) ny; (void) nc; return -1; }
