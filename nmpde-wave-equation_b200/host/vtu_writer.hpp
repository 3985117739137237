// vtu_writer.hpp -- ParaView output of a triangulated scalar field, the file layout the reference
// gets from DataOut::build_patches() + write_vtu_with_pvtu_record (src/WaveEquationBase.cpp:330-365):
// one unstructured-grid piece `<base>_<NNNN>.0.vtu` whose points are the three corners of every
// cell (patches do not share points), every field as point data, and a `<base>_<NNNN>.pvtu` record
// naming the piece.  Arrays are written inline as base64 (uncompressed, UInt32 length header).
#ifndef WAVE_VTU_WRITER_HPP
#define WAVE_VTU_WRITER_HPP

#include <cstdint>
#include <string>
#include <vector>

struct VtuField
{
    std::string name;
    std::vector<double> values; // one per point (3 per cell, in cell order)
};

/// `xyz` holds 3 floats per point and 3 points per cell (cell k = points 3k, 3k+1, 3k+2).
/// Throws std::runtime_error when the file cannot be written.
void write_vtu_piece(const std::string& path, const std::vector<float>& xyz, const std::vector<VtuField>& fields);

/// The parallel record: field declarations and the list of piece files (relative names).
void write_pvtu_record(const std::string& path, const std::vector<std::string>& pieces,
                       const std::vector<std::string>& field_names);

/// deal.II's names: "<base>_<counter, n_digits>.<piece>.vtu" and "<base>_<counter>.pvtu".
std::string vtu_piece_name(const std::string& base, unsigned int counter, unsigned int piece,
                           unsigned int n_digits = 4);
std::string pvtu_record_name(const std::string& base, unsigned int counter, unsigned int n_digits = 4);

#endif
