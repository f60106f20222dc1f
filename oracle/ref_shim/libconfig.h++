/* TEST INFRASTRUCTURE ONLY (oracle/_ref build). The reference's helpers/exception.h:7,34
 * needs a type libconfig::Setting with getPath(); libconfig++ is not installed here. */
#ifndef JB_ORACLE_SHIM_LIBCONFIG_HPP
#define JB_ORACLE_SHIM_LIBCONFIG_HPP
#include <string>
namespace libconfig {
class Setting {
 public:
  std::string getPath() const { return std::string(); }
};
}
#endif
