#include "pcd_io.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

namespace f3ps {
namespace {
bool lzf_decompress(const unsigned char* in, size_t in_len, std::vector<unsigned char>& out, size_t out_len) {
    out.resize(out_len);
    size_t ip = 0, op = 0;
    while (ip < in_len) {
        unsigned ctrl = in[ip++];
        if (ctrl < 32) {
            size_t ln = ctrl + 1;
            if (op + ln > out_len || ip + ln > in_len) return false;
            memcpy(&out[op], &in[ip], ln); ip += ln; op += ln;
        } else {
            size_t ln = ctrl >> 5;
            if (ln == 7) { if (ip >= in_len) return false; ln += in[ip++]; }
            if (ip >= in_len) return false;
            size_t dist = ((ctrl & 0x1f) << 8) + in[ip++] + 1;
            ln += 2;
            if (dist > op || op + ln > out_len) return false;
            size_t ref = op - dist;
            for (size_t k = 0; k < ln; ++k) out[op++] = out[ref++];
        }
    }
    return op == out_len;
}
struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; size_t offset = 0; };
double read_scalar(const unsigned char* p, const Field& f) {
    switch (f.type) {
        case 'F': if (f.size == 4) { float v; memcpy(&v, p, 4); return v; } else { double v; memcpy(&v, p, 8); return v; }
        case 'U': if (f.size == 1) return *p; if (f.size == 2) { uint16_t v; memcpy(&v, p, 2); return v; } { uint32_t v; memcpy(&v, p, 4); return v; }
        default: if (f.size == 1) return *(const int8_t*)p; if (f.size == 2) { int16_t v; memcpy(&v, p, 2); return v; } { int32_t v; memcpy(&v, p, 4); return v; }
    }
}
}

int loadPCDFile(const std::string& path, pcl::PointCloud<pcl::PointXYZRGBL>& cloud) {
    cloud.clear();
    std::ifstream f(path, std::ios::binary);
    if (!f) return -1;
    std::vector<unsigned char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<Field> fields; size_t npts = 0, width = 0, height = 1; std::string mode;
    size_t pos = 0;
    while (pos < raw.size()) {
        size_t end = pos;
        while (end < raw.size() && raw[end] != '\n') ++end;
        std::string line((const char*)&raw[pos], end - pos);
        pos = end + 1;
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line); std::string key; ss >> key;
        if (key == "FIELDS") { std::string n; while (ss >> n) { Field fd; fd.name = n; fields.push_back(fd); } }
        else if (key == "SIZE") { for (auto& fd : fields) ss >> fd.size; }
        else if (key == "TYPE") { for (auto& fd : fields) ss >> fd.type; }
        else if (key == "COUNT") { for (auto& fd : fields) ss >> fd.count; }
        else if (key == "WIDTH") ss >> width;
        else if (key == "HEIGHT") ss >> height;
        else if (key == "POINTS") ss >> npts;
        else if (key == "DATA") { ss >> mode; break; }
    }
    if (!npts) npts = width * height;
    if (fields.empty() || mode.empty()) return -1;
    if (npts > raw.size()) return -1;                      // every encoding spends at least one byte per point
    size_t rec = 0;
    for (auto& fd : fields) { fd.offset = rec; rec += (size_t)fd.size * fd.count; }
    int ix = -1, iy = -1, iz = -1, ic = -1, il = -1;
    for (size_t i = 0; i < fields.size(); ++i) {
        if (fields[i].name == "x") ix = (int)i; else if (fields[i].name == "y") iy = (int)i; else if (fields[i].name == "z") iz = (int)i;
        else if (fields[i].name == "rgb" || fields[i].name == "rgba") ic = (int)i; else if (fields[i].name == "label") il = (int)i;
    }
    if (ix < 0 || iy < 0 || iz < 0) return -1;
    cloud.points.resize(npts); cloud.width = (uint32_t)(width ? width : npts); cloud.height = (uint32_t)height;
    auto colour_of = [&](const unsigned char* p) { uint32_t v; memcpy(&v, p, 4); return v; };
    if (mode == "ascii") {
        std::istringstream ss(std::string((const char*)&raw[std::min(pos, raw.size())], raw.size() - std::min(pos, raw.size())));
        for (size_t i = 0; i < npts; ++i) {
            pcl::PointXYZRGBL& p = cloud.points[i];
            for (size_t k = 0; k < fields.size(); ++k) for (int c = 0; c < fields[k].count; ++c) {
                std::string tok; ss >> tok;
                const bool is_nan = tok == "nan" || tok == "NaN" || tok == "-nan";
                double v = is_nan ? std::numeric_limits<double>::quiet_NaN() : atof(tok.c_str());
                if ((int)k == ix) p.x = (float)v; else if ((int)k == iy) p.y = (float)v; else if ((int)k == iz) p.z = (float)v;
                else if ((int)k == ic) {
                    // PCL >= 1.8 writes an ascii rgb field of TYPE F as the uint32 reinterpretation of the packed colour
                    // (e.g. 4285098345); older files carry the float's decimal form
                    const bool as_int = !is_nan && tok.find_first_of(".eEnN") == std::string::npos;
                    if (fields[k].type == 'F' && !as_int) { float fv = (float)v; memcpy(&p.rgba, &fv, 4); }
                    else p.rgba = (uint32_t)strtoul(tok.c_str(), nullptr, 10);
                }
                else if ((int)k == il) p.label = (uint32_t)v;
            }
        }
    } else if (mode == "binary") {
        if (pos + rec * npts > raw.size()) return -1;
        for (size_t i = 0; i < npts; ++i) {
            const unsigned char* r = &raw[pos + i * rec]; pcl::PointXYZRGBL& p = cloud.points[i];
            p.x = (float)read_scalar(r + fields[ix].offset, fields[ix]); p.y = (float)read_scalar(r + fields[iy].offset, fields[iy]);
            p.z = (float)read_scalar(r + fields[iz].offset, fields[iz]);
            if (ic >= 0) p.rgba = colour_of(r + fields[ic].offset);
            if (il >= 0) p.label = (uint32_t)read_scalar(r + fields[il].offset, fields[il]);
        }
    } else if (mode == "binary_compressed") {
        if (pos + 8 > raw.size()) return -1;
        uint32_t csz, usz; memcpy(&csz, &raw[pos], 4); memcpy(&usz, &raw[pos + 4], 4);
        if (pos + 8 + csz > raw.size()) return -1;
        std::vector<unsigned char> data;
        if (!lzf_decompress(&raw[pos + 8], csz, data, usz)) return -1;
        std::vector<size_t> col(fields.size()); size_t o = 0;           // structure of arrays: all x, all y, ...
        for (size_t k = 0; k < fields.size(); ++k) { col[k] = o; o += (size_t)fields[k].size * fields[k].count * npts; }
        if (o > data.size()) return -1;
        for (size_t i = 0; i < npts; ++i) {
            pcl::PointXYZRGBL& p = cloud.points[i];
            p.x = (float)read_scalar(&data[col[ix] + i * fields[ix].size], fields[ix]);
            p.y = (float)read_scalar(&data[col[iy] + i * fields[iy].size], fields[iy]);
            p.z = (float)read_scalar(&data[col[iz] + i * fields[iz].size], fields[iz]);
            if (ic >= 0) p.rgba = colour_of(&data[col[ic] + i * 4]);
            if (il >= 0) p.label = (uint32_t)read_scalar(&data[col[il] + i * fields[il].size], fields[il]);
        }
    } else return -1;
    return 0;
}

int savePCDFileASCII(const std::string& path, const pcl::PointCloud<pcl::PointXYZL>& cloud) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return -1;
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z label\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
               "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA ascii\n", cloud.size(), cloud.size());
    for (const auto& p : cloud.points) fprintf(f, "%.9g %.9g %.9g %u\n", p.x, p.y, p.z, p.label);
    fclose(f);
    return 0;
}
} // namespace f3ps
