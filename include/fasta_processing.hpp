/**
 * fasta_processing.hpp -- drop-in for the reference's FASTA ingest (src/fasta_processing.hpp:18-23).
 * Host only.  Same record and split rules as src/fasta_processing.cpp:79-211: records
 * split at blank lines, a line holding a space drops its record, every non-ACGT byte (incl. the '\r'
 * of CRLF files) starts a new string, lower case accepted, unreadable file => stderr + exit(1).
 * The set builders of kmer.hpp do not go through these byte-per-base strings: they use the 2-bit
 * packer + segment table of the C ABI (sks_fasta_parse_file) and sketch on the device.
 */
#ifndef SKS_FASTA_PROCESSING_HPP
#define SKS_FASTA_PROCESSING_HPP
#include <cstdint>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "logging.hpp"

/** One byte per base: A/a 0, C/c 1, G/g 2, T/t 3 (src/fasta_processing.cpp:35-69). */
typedef std::vector<uint8_t> acgt_string;

std::vector<std::string> strings_from_fasta(const char fasta_filename[]);
void add_nucleotide_strings(std::vector<acgt_string> &return_strings, const std::string &raw_string);
std::vector<acgt_string> cut_nucleotide_strings(const std::vector<std::string> &raw_strings);
std::vector<acgt_string> nucleotide_strings_from_fasta_file(const char fasta_filename[]);
#endif
