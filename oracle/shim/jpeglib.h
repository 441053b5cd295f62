/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
 * Declaration-only stand-in for libjpeg: the reference's Image.h hard-defines cimg_use_jpeg, so CImg.h includes
 * <jpeglib.h> and instantiates its JPEG reader/writer.  This image has no libjpeg; the oracle feeds the reference
 * PPM files (CImg's own PNM reader), so none of these entry points is ever reached - oracle/ref_jpeg_stub.c makes
 * them abort loudly if they are. */
#ifndef HPMVS_ORACLE_JPEGLIB_SHIM_H
#define HPMVS_ORACLE_JPEGLIB_SHIM_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
#define JMSG_LENGTH_MAX 200
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define METHODDEF(type) static type
typedef int boolean;
typedef unsigned int JDIMENSION;
typedef unsigned char JSAMPLE;
typedef JSAMPLE* JSAMPROW;
typedef JSAMPROW* JSAMPARRAY;
typedef enum { JCS_UNKNOWN, JCS_GRAYSCALE, JCS_RGB, JCS_YCbCr, JCS_CMYK, JCS_YCCK } J_COLOR_SPACE;
struct jpeg_common_struct;
typedef struct jpeg_common_struct* j_common_ptr;
struct jpeg_error_mgr {
    void (*error_exit)(j_common_ptr cinfo);
    void (*format_message)(j_common_ptr cinfo, char* buffer);
};
struct jpeg_common_struct { struct jpeg_error_mgr* err; };
struct jpeg_decompress_struct {
    struct jpeg_error_mgr* err;
    JDIMENSION output_width, output_height, output_scanline;
    int output_components;
};
struct jpeg_compress_struct {
    struct jpeg_error_mgr* err;
    JDIMENSION image_width, image_height, next_scanline;
    int input_components;
    J_COLOR_SPACE in_color_space;
};
typedef struct jpeg_decompress_struct* j_decompress_ptr;
typedef struct jpeg_compress_struct* j_compress_ptr;
struct jpeg_error_mgr* jpeg_std_error(struct jpeg_error_mgr* err);
void jpeg_destroy(j_common_ptr cinfo);
void jpeg_create_decompress(j_decompress_ptr cinfo);
void jpeg_stdio_src(j_decompress_ptr cinfo, FILE* f);
int jpeg_read_header(j_decompress_ptr cinfo, boolean require_image);
boolean jpeg_start_decompress(j_decompress_ptr cinfo);
JDIMENSION jpeg_read_scanlines(j_decompress_ptr cinfo, JSAMPARRAY scanlines, JDIMENSION max_lines);
boolean jpeg_finish_decompress(j_decompress_ptr cinfo);
void jpeg_destroy_decompress(j_decompress_ptr cinfo);
void jpeg_create_compress(j_compress_ptr cinfo);
void jpeg_stdio_dest(j_compress_ptr cinfo, FILE* f);
void jpeg_set_defaults(j_compress_ptr cinfo);
void jpeg_set_quality(j_compress_ptr cinfo, int quality, boolean force_baseline);
void jpeg_start_compress(j_compress_ptr cinfo, boolean write_all_tables);
JDIMENSION jpeg_write_scanlines(j_compress_ptr cinfo, JSAMPARRAY scanlines, JDIMENSION num_lines);
void jpeg_finish_compress(j_compress_ptr cinfo);
void jpeg_destroy_compress(j_compress_ptr cinfo);
#ifdef __cplusplus
}
#endif
#endif
