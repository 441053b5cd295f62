// File formats on either side of the path ("next" row f-4), host C++ only:
//   * NVM_V3 reader  - replaces mo3d::NVMReader::readFile (/root/reference/src/hpmvs/NVMReader.cpp:31-155)
//   * ext-PLY writer - replaces DynOctTree::toExtPly      (/root/reference/include/hpmvs/doctree.h:525-622)
//   * binary PPM (P6) level-0 image reader: the reference decodes JPEG through CImg/libjpeg (Image.cpp:46); this image
//     has no JPEG decoder, so scenes carry PPM files instead (documented deviation).
#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hpmvs_b200.h"
#include "undistort_math.h"

struct hpmvs_nvm {
    struct Cam { std::string filename; double f, q[4], c[3], r; };
    std::vector<Cam> cams;
    std::vector<double> xyz, rgb, meas_xy;
    std::vector<int32_t> offsets, meas_cam, meas_feat;
    int n_models = 0;
};

namespace {
std::string folder_of(const std::string& path) {
    const size_t p = path.find_last_of('/');
    return p == std::string::npos ? std::string("") : path.substr(0, p);
}
}  // namespace

extern "C" {

int hpmvs_nvm_open(const char* path, int fix_path, hpmvs_nvm_t** out) {
    if (!path || !out) return HPMVS_E_ARG;
    *out = nullptr;
    std::ifstream in(path);
    if (!in.good()) return HPMVS_E_ARG;
    std::string header;
    in >> header;
    if (strcasecmp("NVM_V3", header.c_str()) != 0) return HPMVS_E_ARG;     // NVMReader.cpp:129
    hpmvs_nvm* m = new hpmvs_nvm;
    m->offsets.push_back(0);
    const std::string folder = folder_of(path);
    // models follow each other until one with 0 cameras (NVMReader.cpp:134-150); only models[0] is used
    // downstream (src/main.cpp:112-116) but the stream must be consumed the same way
    bool first = true;
    while (in.good()) {
        int ncam = 0;
        if (!(in >> ncam) || ncam <= 0) break;
        std::vector<hpmvs_nvm::Cam> cams(ncam);
        for (auto& c : cams) {
            int check = 0;
            in >> c.filename >> c.f >> c.q[0] >> c.q[1] >> c.q[2] >> c.q[3] >> c.c[0] >> c.c[1] >> c.c[2] >> c.r >> check;
            std::replace(c.filename.begin(), c.filename.end(), '"', ' ');            // NVMReader.cpp:73
            if (fix_path && !c.filename.empty() && c.filename[0] != '/') c.filename = folder.empty() ? c.filename : folder + "/" + c.filename;
        }
        int npts = 0;
        in >> npts;
        m->n_models++;
        for (int i = 0; i < npts; i++) {
            double xyz[3], rgb[3];
            int nm = 0;
            in >> xyz[0] >> xyz[1] >> xyz[2] >> rgb[0] >> rgb[1] >> rgb[2] >> nm;
            if (!in.good() && !in.eof()) { delete m; return HPMVS_E_ARG; }
            for (int k = 0; k < nm; k++) {
                int img = 0, feat = 0; double u = 0, v = 0;
                in >> img >> feat >> u >> v;
                if (first) { m->meas_cam.push_back(img); m->meas_feat.push_back(feat); m->meas_xy.push_back(u); m->meas_xy.push_back(v); }
            }
            if (first) {
                m->xyz.insert(m->xyz.end(), xyz, xyz + 3); m->rgb.insert(m->rgb.end(), rgb, rgb + 3);
                m->offsets.push_back((int32_t)m->meas_cam.size());
            }
        }
        if (first) m->cams = cams;
        first = false;
    }
    *out = m;
    return 0;
}

void hpmvs_nvm_close(hpmvs_nvm_t* m) { delete m; }
int hpmvs_nvm_num_models(const hpmvs_nvm_t* m) { return m ? m->n_models : 0; }
int hpmvs_nvm_num_cameras(const hpmvs_nvm_t* m) { return m ? (int)m->cams.size() : 0; }
int hpmvs_nvm_num_points(const hpmvs_nvm_t* m) { return m ? (int)m->offsets.size() - 1 : 0; }
int hpmvs_nvm_num_measurements(const hpmvs_nvm_t* m) { return m ? (int)m->meas_cam.size() : 0; }

int hpmvs_nvm_camera(const hpmvs_nvm_t* m, int i, char* filename, int cap, double* f, double q_wxyz[4], double c[3], double* r) {
    if (!m || i < 0 || i >= (int)m->cams.size()) return HPMVS_E_ARG;
    const auto& cam = m->cams[i];
    if (filename && cap > 0) { strncpy(filename, cam.filename.c_str(), cap - 1); filename[cap - 1] = 0; }
    if (f) *f = cam.f;
    if (q_wxyz) memcpy(q_wxyz, cam.q, sizeof(cam.q));
    if (c) memcpy(c, cam.c, sizeof(cam.c));
    if (r) *r = cam.r;
    return 0;
}

int hpmvs_nvm_points(const hpmvs_nvm_t* m, double* xyz, double* rgb, int32_t* offsets, int32_t* meas_cam, int32_t* meas_feat, double* meas_xy) {
    if (!m) return HPMVS_E_ARG;
    if (xyz) std::copy(m->xyz.begin(), m->xyz.end(), xyz);
    if (rgb) std::copy(m->rgb.begin(), m->rgb.end(), rgb);
    if (offsets) std::copy(m->offsets.begin(), m->offsets.end(), offsets);
    if (meas_cam) std::copy(m->meas_cam.begin(), m->meas_cam.end(), meas_cam);
    if (meas_feat) std::copy(m->meas_feat.begin(), m->meas_feat.end(), meas_feat);
    if (meas_xy) std::copy(m->meas_xy.begin(), m->meas_xy.end(), meas_xy);
    return 0;
}

// NVMReader::saveNVM (src/hpmvs/NVMReader.cpp:157-183) for the model that was read: 12 significant digits, the stream operators of
// NVM_Model / NVM_Camera / NVM_Point / NVM_Measurement (:34-111: colours as integers, a blank before every measurement field), the
// terminating empty model "0" without a newline
int hpmvs_nvm_write(const hpmvs_nvm_t* m, const char* path) {
    if (!m || !path) return HPMVS_E_ARG;
    std::ofstream out(path);
    if (!out.good()) return HPMVS_E_ARG;
    out << std::setprecision(12);
    out << "NVM_V3" << std::endl;
    const int ncam = (int)m->cams.size(), npts = (int)m->offsets.size() - 1;
    out << std::endl << ncam << std::endl;
    for (const auto& c : m->cams) {
        out << c.filename << " " << c.f << " " << c.q[0] << " " << c.q[1] << " " << c.q[2] << " " << c.q[3] << " "
            << c.c[0] << " " << c.c[1] << " " << c.c[2] << " " << c.r << " " << 0 << std::endl;
    }
    if (ncam > 0) out << std::endl << npts << std::endl;
    for (int i = 0; i < npts; i++) {
        out << m->xyz[3 * i] << " " << m->xyz[3 * i + 1] << " " << m->xyz[3 * i + 2] << " " << (int)m->rgb[3 * i] << " "
            << (int)m->rgb[3 * i + 1] << " " << (int)m->rgb[3 * i + 2];
        const int a = m->offsets[i], b = m->offsets[i + 1];
        out << " " << (b - a);
        for (int k = a; k < b; k++)
            out << " " << m->meas_cam[k] << " " << m->meas_feat[k] << " " << m->meas_xy[2 * k] << " " << m->meas_xy[2 * k + 1];
        out << std::endl;
    }
    out << "0";
    return out.good() ? 0 : HPMVS_E_ARG;
}

// binary PPM; returns the size with rgb == NULL, fills rgb (3*w*h bytes, interleaved) otherwise
int hpmvs_ppm_read(const char* path, int* width, int* height, uint8_t* rgb) {
    if (!path || !width || !height) return HPMVS_E_ARG;
    FILE* fh = fopen(path, "rb");
    if (!fh) return HPMVS_E_ARG;
    int vals[3], got = 0, c;
    char magic[3] = {0, 0, 0};
    if (fread(magic, 1, 2, fh) != 2 || magic[0] != 'P' || magic[1] != '6') { fclose(fh); return HPMVS_E_ARG; }
    while (got < 3 && (c = fgetc(fh)) != EOF) {
        if (c == '#') { while ((c = fgetc(fh)) != EOF && c != '\n') {} continue; }
        if (isspace(c)) continue;
        int v = 0;
        while (c != EOF && isdigit(c)) { v = v * 10 + (c - '0'); c = fgetc(fh); }
        vals[got++] = v;     // the single whitespace after maxval has just been consumed
    }
    if (got < 3 || vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0) { fclose(fh); return HPMVS_E_ARG; }
    *width = vals[0]; *height = vals[1];
    int rc = 0;
    if (rgb) {
        const size_t n = (size_t)3 * vals[0] * vals[1];
        if (fread(rgb, 1, n, fh) != n) rc = HPMVS_E_ARG;
    }
    fclose(fh);
    return rc;
}

// Image::undistort (src/hpmvs/Image.cpp:68-149): VisualSFM's one-parameter radial model undone on the level-0 image.
// For every target pixel the distorted source position is found in closed form (Cardano, double / complex<double> exactly as
// the reference writes it), sampled with CImg's _linear_atXY (thirdLibs/cimg/CImg.h:12218-12235, f32) when it lies strictly
// inside the 1-pixel border, and the f32 result is truncated to u8 when the float image is assigned back (Image.cpp:143-144).
// Pixels whose source falls outside stay 0 (the reference leaves its freshly allocated float buffer untouched there, Q13).
// Host code on purpose: pow / sqrt / complex pow come from the same libm the reference would use.
int hpmvs_undistort_rgb(const uint8_t* rgb, int width, int height, double f_in, double k1_in, uint8_t* out, uint8_t* written) {
    if (!rgb || !out || width <= 0 || height <= 0) return HPMVS_E_ARG;
    const float f_ = (float)f_in, k1_ = (float)k1_in;           // Image::init stores them as float (Image.cpp:36-37, Image.h:78)
    const size_t npx = (size_t)width * height;
    if (k1_ == 0) { memcpy(out, rgb, 3 * npx); if (written) memset(written, 1, npx); return 0; }   // Image::load only calls undistort() when k1_ != 0 (Image.cpp:51)
    memset(out, 0, 3 * npx);
    if (written) memset(written, 0, npx);
    auto at = [&](unsigned x, unsigned y, int c) -> float { return (float)rgb[3 * ((size_t)y * width + x) + c]; };
    for (int ix = 0; ix < width; ix++) {
        for (int iy = 0; iy < height; iy++) {
            float y = (float)(iy - height / 2.0);
            float x = (float)(ix - width / 2.0);
            x /= f_;
            y /= f_;
            if (y == 0) y = 1e-3;
            // source position in the distorted image (undistort_math.h: Cardano's root, the reference's operations in its order)
            const double kr = k1_ * ((double)(y * y) + (double)(x * x));
            const ud::Source src = (k1_ > 0) ? ud::source_positive_k1(x, y, kr) : ud::source_negative_k1_host(x, y, kr);
            const float mx = src.mx, my = src.my;
            x = mx * (float)f_ + width / 2.0f;
            y = my * (float)f_ + height / 2.0f;
            if (x > 1 && x < width - 1 && y > 1 && y < height - 1) {
                // CImg<unsigned char>::_linear_atXY(x, y, 0, c)
                const unsigned W = (unsigned)width, H = (unsigned)height;
                const float nfx = x < 0 ? 0 : (x > W - 1 ? W - 1 : x), nfy = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
                const unsigned px = (unsigned)nfx, py = (unsigned)nfy;
                const float dx = nfx - px, dy = nfy - py;
                const unsigned nx = dx > 0 ? px + 1 : px, ny = dy > 0 ? py + 1 : py;
                for (int c = 0; c < 3; c++) {
                    const float Icc = at(px, py, c), Inc = at(nx, py, c), Icn = at(px, ny, c), Inn = at(nx, ny, c);
                    const float v = Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
                    out[3 * ((size_t)iy * width + ix) + c] = (uint8_t)v;
                }
                if (written) written[(size_t)iy * width + ix] = 1;
            }
        }
    }
    return 0;
}

int hpmvs_ply_write_ext(const char* path, int n, const hpmvs_patch_t* p, int binary, int normal, int scale, int visibility) {
    if (!path || n < 0 || (n > 0 && !p)) return HPMVS_E_ARG;
    {
        std::ofstream f(path, std::ofstream::out);
        if (!f.good()) return HPMVS_E_ARG;
        f << "ply" << std::endl;
        if (binary) f << "format binary_little_endian 1.0" << std::endl;
        else f << "format ascii 1.0" << std::endl;
        f << "element vertex " << n << std::endl;
        f << "property float x" << std::endl << "property float y" << std::endl << "property float z" << std::endl;
        if (normal) f << "property float nx" << std::endl << "property float ny" << std::endl << "property float nz" << std::endl;
        f << "property uchar red" << std::endl << "property uchar green" << std::endl << "property uchar blue" << std::endl;
        if (scale) f << "property float scalar_scale" << std::endl;
        if (visibility) {
            f << "element point_visibility " << n << std::endl;
            f << "property list uint uint visible_cameras" << std::endl;
        }
        f << "end_header" << std::endl;
    }
    std::ofstream d(path, binary ? std::ofstream::binary | std::ofstream::app : std::ofstream::app);
    for (int i = 0; i < n; i++) {
        const hpmvs_patch_t& q = p[i];
        if (binary) {
            d.write((const char*)q.center, 3 * sizeof(float));
            if (normal) d.write((const char*)q.normal, 3 * sizeof(float));
            const unsigned char c[3] = {(unsigned char)q.color[0], (unsigned char)q.color[1], (unsigned char)q.color[2]};
            d.write((const char*)c, 3);
            if (scale) d.write((const char*)&q.scale, sizeof(float));
        } else {
            d << q.center[0] << " " << q.center[1] << " " << q.center[2] << " ";
            if (normal) d << q.normal[0] << " " << q.normal[1] << " " << q.normal[2] << " ";
            d << (int)q.color[0] << " " << (int)q.color[1] << " " << (int)q.color[2] << " ";
            if (scale) d << q.scale << " ";
            d << std::endl;
        }
    }
    if (visibility)
        for (int i = 0; i < n; i++) {
            const hpmvs_patch_t& q = p[i];
            if (binary) {
                const uint32_t k = (uint32_t)q.nimages;
                d.write((const char*)&k, sizeof(uint32_t));
                for (int j = 0; j < q.nimages; j++) { const uint32_t id = (uint32_t)q.images[j]; d.write((const char*)&id, sizeof(uint32_t)); }
            } else {
                d << (int)q.nimages << " ";
                for (int j = 0; j < q.nimages; j++) d << (uint32_t)q.images[j] << " ";
                d << std::endl;
            }
        }
    d.flush();
    return d.good() ? 0 : HPMVS_E_ARG;
}

}  // extern "C"
