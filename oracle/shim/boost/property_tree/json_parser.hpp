// See ptree.hpp: the JSON reader / writer are named by include/json.h but never called on this path.
#ifndef VISGEOM_ORACLE_JSON_PARSER_SHIM
#define VISGEOM_ORACLE_JSON_PARSER_SHIM
#include "ptree.hpp"
namespace boost { namespace property_tree {
inline void read_json(const std::string &, ptree &) { throw std::runtime_error("read_json: not available in the oracle build"); }
inline void write_json(const std::string &, const ptree &) { throw std::runtime_error("write_json: not available in the oracle build"); }
}}
#endif
