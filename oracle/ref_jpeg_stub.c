/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.  See oracle/shim/jpeglib.h: there is no libjpeg in this image and
 * the oracle only ever hands the reference PPM images, so every libjpeg entry point aborts. */
#include "jpeglib.h"
#include <stdlib.h>
static void die(const char* f) { fprintf(stderr, "oracle/_ref: %s called, but this build has no JPEG codec (use .ppm images)\n", f); abort(); }
struct jpeg_error_mgr* jpeg_std_error(struct jpeg_error_mgr* err) { (void)err; die("jpeg_std_error"); return 0; }
void jpeg_destroy(j_common_ptr c) { (void)c; die("jpeg_destroy"); }
void jpeg_create_decompress(j_decompress_ptr c) { (void)c; die("jpeg_create_decompress"); }
void jpeg_stdio_src(j_decompress_ptr c, FILE* f) { (void)c; (void)f; die("jpeg_stdio_src"); }
int jpeg_read_header(j_decompress_ptr c, boolean r) { (void)c; (void)r; die("jpeg_read_header"); return 0; }
boolean jpeg_start_decompress(j_decompress_ptr c) { (void)c; die("jpeg_start_decompress"); return 0; }
JDIMENSION jpeg_read_scanlines(j_decompress_ptr c, JSAMPARRAY s, JDIMENSION m) { (void)c; (void)s; (void)m; die("jpeg_read_scanlines"); return 0; }
boolean jpeg_finish_decompress(j_decompress_ptr c) { (void)c; die("jpeg_finish_decompress"); return 0; }
void jpeg_destroy_decompress(j_decompress_ptr c) { (void)c; die("jpeg_destroy_decompress"); }
void jpeg_create_compress(j_compress_ptr c) { (void)c; die("jpeg_create_compress"); }
void jpeg_stdio_dest(j_compress_ptr c, FILE* f) { (void)c; (void)f; die("jpeg_stdio_dest"); }
void jpeg_set_defaults(j_compress_ptr c) { (void)c; die("jpeg_set_defaults"); }
void jpeg_set_quality(j_compress_ptr c, int q, boolean b) { (void)c; (void)q; (void)b; die("jpeg_set_quality"); }
void jpeg_start_compress(j_compress_ptr c, boolean w) { (void)c; (void)w; die("jpeg_start_compress"); }
JDIMENSION jpeg_write_scanlines(j_compress_ptr c, JSAMPARRAY s, JDIMENSION n) { (void)c; (void)s; (void)n; die("jpeg_write_scanlines"); return 0; }
void jpeg_finish_compress(j_compress_ptr c) { (void)c; die("jpeg_finish_compress"); }
void jpeg_destroy_compress(j_compress_ptr c) { (void)c; die("jpeg_destroy_compress"); }
