// vtu_writer.cpp -- see vtu_writer.hpp.
#include "vtu_writer.hpp"

#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>

namespace
{
// base64 of a UInt32 byte count followed by the raw array (VTK "binary" inline encoding,
// header_type UInt32, no compressor)
std::string encoded(const void* data, const size_t n_bytes)
{
    if (n_bytes > 0xffffffffull)
        throw std::runtime_error("VTU array larger than 4 GiB: raise the piece count");
    std::string raw(sizeof(uint32_t) + n_bytes, '\0');
    const uint32_t header = static_cast<uint32_t>(n_bytes);
    std::memcpy(&raw[0], &header, sizeof header);
    if (n_bytes)
        std::memcpy(&raw[sizeof header], data, n_bytes);

    static const char digits[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string out;
    out.reserve((raw.size() + 2) / 3 * 4);
    size_t k = 0;
    for (; k + 2 < raw.size(); k += 3)
    {
        const uint32_t w = (uint32_t(uint8_t(raw[k])) << 16) | (uint32_t(uint8_t(raw[k + 1])) << 8) | uint8_t(raw[k + 2]);
        out += digits[(w >> 18) & 63];
        out += digits[(w >> 12) & 63];
        out += digits[(w >> 6) & 63];
        out += digits[w & 63];
    }
    if (k < raw.size())
    {
        const bool two = k + 1 < raw.size();
        const uint32_t w = (uint32_t(uint8_t(raw[k])) << 16) | (two ? uint32_t(uint8_t(raw[k + 1])) << 8 : 0u);
        out += digits[(w >> 18) & 63];
        out += digits[(w >> 12) & 63];
        out += two ? digits[(w >> 6) & 63] : '=';
        out += '=';
    }
    return out;
}

template <class T>
void data_array(std::ostream& os, const char* type, const std::string& name, const int components,
                const std::vector<T>& v)
{
    os << "        <DataArray type=\"" << type << "\"";
    if (!name.empty())
        os << " Name=\"" << name << "\"";
    if (components > 1)
        os << " NumberOfComponents=\"" << components << "\"";
    os << " format=\"binary\">\n          " << encoded(v.data(), v.size() * sizeof(T)) << "\n        </DataArray>\n";
}

std::string counter_text(unsigned int counter, unsigned int n_digits)
{
    std::ostringstream s;
    s << std::setw(static_cast<int>(n_digits)) << std::setfill('0') << counter;
    return s.str();
}
} // namespace

std::string vtu_piece_name(const std::string& base, unsigned int counter, unsigned int piece, unsigned int n_digits)
{
    return base + "_" + counter_text(counter, n_digits) + "." + std::to_string(piece) + ".vtu";
}

std::string pvtu_record_name(const std::string& base, unsigned int counter, unsigned int n_digits)
{
    return base + "_" + counter_text(counter, n_digits) + ".pvtu";
}

void write_vtu_piece(const std::string& path, const std::vector<float>& xyz, const std::vector<VtuField>& fields)
{
    if (xyz.size() % 9 != 0)
        throw std::invalid_argument("write_vtu_piece: 3 points of 3 coordinates per cell expected");
    const size_t n_points = xyz.size() / 3, n_cells = n_points / 3;
    for (const VtuField& f : fields)
        if (f.values.size() != n_points)
            throw std::invalid_argument("write_vtu_piece: field '" + f.name + "' does not have one value per point");

    std::vector<int32_t> connectivity(n_points), offsets(n_cells);
    std::vector<uint8_t> types(n_cells, 5); // VTK_TRIANGLE
    for (size_t p = 0; p < n_points; ++p)
        connectivity[p] = static_cast<int32_t>(p);
    for (size_t c = 0; c < n_cells; ++c)
        offsets[c] = static_cast<int32_t>(3 * (c + 1));

    std::ofstream os(path, std::ios::binary);
    if (!os)
        throw std::runtime_error("cannot open " + path);
    os << "<?xml version=\"1.0\"?>\n"
       << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt32\">\n"
       << "  <UnstructuredGrid>\n"
       << "    <Piece NumberOfPoints=\"" << n_points << "\" NumberOfCells=\"" << n_cells << "\">\n"
       << "      <Points>\n";
    data_array(os, "Float32", "", 3, xyz);
    os << "      </Points>\n      <Cells>\n";
    data_array(os, "Int32", "connectivity", 1, connectivity);
    data_array(os, "Int32", "offsets", 1, offsets);
    data_array(os, "UInt8", "types", 1, types);
    os << "      </Cells>\n      <PointData" << (fields.empty() ? "" : " Scalars=\"" + fields.front().name + "\"") << ">\n";
    for (const VtuField& f : fields)
        data_array(os, "Float64", f.name, 1, f.values);
    os << "      </PointData>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n";
    if (!os)
        throw std::runtime_error("write failed: " + path);
}

void write_pvtu_record(const std::string& path, const std::vector<std::string>& pieces,
                       const std::vector<std::string>& field_names)
{
    std::ofstream os(path);
    if (!os)
        throw std::runtime_error("cannot open " + path);
    os << "<?xml version=\"1.0\"?>\n"
       << "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
       << "  <PUnstructuredGrid GhostLevel=\"0\">\n"
       << "    <PPointData" << (field_names.empty() ? "" : " Scalars=\"" + field_names.front() + "\"") << ">\n";
    for (const std::string& name : field_names)
        os << "      <PDataArray type=\"Float64\" Name=\"" << name << "\" format=\"binary\"/>\n";
    os << "    </PPointData>\n"
       << "    <PPoints>\n      <PDataArray type=\"Float32\" NumberOfComponents=\"3\"/>\n    </PPoints>\n";
    for (const std::string& piece : pieces)
        os << "    <Piece Source=\"" << piece << "\"/>\n";
    os << "  </PUnstructuredGrid>\n</VTKFile>\n";
    if (!os)
        throw std::runtime_error("write failed: " + path);
}
