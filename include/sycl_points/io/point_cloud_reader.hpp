// PointCloudReader::readFile — I/io/point_cloud_reader.hpp:494-548, PLY only (ASCII and
// binary_little_endian vertex records of scalar properties); what the bundled cpp/data/*.ply and
// typical LiDAR exports use.  PCD and the writer are outside the registration path.
#pragma once

#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {

class PointCloudReader {
    struct Property {
        std::string name;
        size_t bytes;
        char kind;  // 'f' float32, 'd' float64, 'i' signed, 'u' unsigned
    };
    static Property parse_property(const std::string& type, const std::string& name) {
        if (type == "float" || type == "float32") return {name, 4, 'f'};
        if (type == "double" || type == "float64") return {name, 8, 'd'};
        if (type == "char" || type == "int8") return {name, 1, 'i'};
        if (type == "uchar" || type == "uint8") return {name, 1, 'u'};
        if (type == "short" || type == "int16") return {name, 2, 'i'};
        if (type == "ushort" || type == "uint16") return {name, 2, 'u'};
        if (type == "int" || type == "int32") return {name, 4, 'i'};
        if (type == "uint" || type == "uint32") return {name, 4, 'u'};
        throw std::runtime_error("[PointCloudReader::readFile] unsupported PLY property type: " + type);
    }
    static double decode(const unsigned char* p, const Property& pr) {
        switch (pr.kind) {
            case 'f': { float v; std::memcpy(&v, p, 4); return v; }
            case 'd': { double v; std::memcpy(&v, p, 8); return v; }
            case 'i': {
                if (pr.bytes == 1) { int8_t v; std::memcpy(&v, p, 1); return v; }
                if (pr.bytes == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
                int32_t v; std::memcpy(&v, p, 4); return v;
            }
            default: {
                if (pr.bytes == 1) { uint8_t v; std::memcpy(&v, p, 1); return v; }
                if (pr.bytes == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; }
                uint32_t v; std::memcpy(&v, p, 4); return v;
            }
        }
    }

public:
    static PointCloudCPU readFile(const std::string& filename, bool /*read_rgb*/ = true, bool read_intensity = true) {
        std::ifstream file(filename, std::ios::binary);
        if (!file.is_open()) throw std::runtime_error("[PointCloudReader::readFile] Failed to open file: " + filename);
        std::string line;
        std::getline(file, line);
        if (line.rfind("ply", 0) != 0)
            throw std::runtime_error("[PointCloudReader::readFile] only PLY files are supported: " + filename);
        bool binary = false, in_vertex = false;
        size_t count = 0;
        std::vector<Property> props;
        while (std::getline(file, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::istringstream ss(line);
            std::string tok;
            ss >> tok;
            if (tok == "format") {
                ss >> tok;
                if (tok == "binary_little_endian") binary = true;
                else if (tok != "ascii") throw std::runtime_error("[PointCloudReader::readFile] unsupported PLY format: " + tok);
            } else if (tok == "element") {
                std::string name;
                ss >> name;
                in_vertex = (name == "vertex");
                if (in_vertex) ss >> count;
            } else if (tok == "property" && in_vertex) {
                std::string type, name;
                ss >> type >> name;
                props.push_back(parse_property(type, name));
            } else if (tok == "end_header") {
                break;
            }
        }
        int ix = -1, iy = -1, iz = -1, ii = -1;
        size_t stride = 0;
        std::vector<size_t> offset(props.size());
        for (size_t p = 0; p < props.size(); ++p) {
            offset[p] = stride;
            stride += props[p].bytes;
            if (props[p].name == "x") ix = (int)p;
            if (props[p].name == "y") iy = (int)p;
            if (props[p].name == "z") iz = (int)p;
            if (props[p].name == "intensity" || props[p].name == "scalar_intensity") ii = (int)p;
        }
        if (ix < 0 || iy < 0 || iz < 0) throw std::runtime_error("[PointCloudReader::readFile] PLY has no x/y/z");
        PointCloudCPU cloud;
        cloud.points->resize(count);
        const bool want_i = read_intensity && ii >= 0;
        if (want_i) cloud.intensities->resize(count);
        if (binary) {
            std::vector<unsigned char> buf(count * stride);
            file.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
            if ((size_t)file.gcount() != buf.size()) throw std::runtime_error("[PointCloudReader::readFile] truncated PLY");
            for (size_t i = 0; i < count; ++i) {
                const unsigned char* r = buf.data() + i * stride;
                (*cloud.points)[i] = PointType((float)decode(r + offset[ix], props[ix]), (float)decode(r + offset[iy], props[iy]),
                                               (float)decode(r + offset[iz], props[iz]), 1.0f);
                if (want_i) (*cloud.intensities)[i] = (float)decode(r + offset[ii], props[ii]);
            }
        } else {
            std::vector<double> v(props.size());
            for (size_t i = 0; i < count; ++i) {
                for (auto& x : v) file >> x;
                (*cloud.points)[i] = PointType((float)v[ix], (float)v[iy], (float)v[iz], 1.0f);
                if (want_i) (*cloud.intensities)[i] = (float)v[ii];
            }
        }
        return cloud;
    }

    static PointCloudShared readFile(const std::string& filename, const sycl_utils::DeviceQueue& queue,
                                     bool read_rgb = true, bool read_intensity = true) {
        return PointCloudShared(queue, readFile(filename, read_rgb, read_intensity));
    }
    static PointCloudShared readFile(const sycl_utils::DeviceQueue& queue, const std::string& filename,
                                     bool read_rgb = true, bool read_intensity = true) {
        return readFile(filename, queue, read_rgb, read_intensity);
    }
};

}  // namespace sycl_points
