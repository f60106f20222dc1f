// config.h — reader for JAMS configuration files (libconfig grammar) and the left-to-right patch merge JAMS
// applies to its command-line arguments.
//
// Replaces, for this repository's host layer, libconfig++ v1.8.1 as JAMS uses it (reference CMakeLists.txt:63,
// helpers/defaults.h:16-19: options AutoConvert | AllowScientificNotation | SemicolonSeparators):
//   * grammar: `name = value` or `name : value`, terminated by `;`, `,` or nothing; groups `{ }`, lists `( )`,
//     arrays `[ ]`; scalars: booleans (true/false, any case), integers (decimal / 0x hex, optional L suffix),
//     floats (incl. scientific notation), strings ("..." with escapes; adjacent strings concatenate);
//     comments `//`, `#`, `/* */`
//   * lookup by dotted path (`solver.t_step`), by name and by index; AutoConvert: an integer setting can be read
//     as a double and vice versa
//   * merge: core/jams++.cc:48-84 + interface/config.cc:12-145 — every command-line argument is a file name or a
//     config string; later arguments overwrite scalars, add missing settings, and patch aggregates element by
//     element (lists/arrays by position, groups by name)
#ifndef JAMS_B200_HOST_CONFIG_H
#define JAMS_B200_HOST_CONFIG_H

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace jams_b200 {

class ConfigError : public std::runtime_error {
 public:
  explicit ConfigError(const std::string &what) : std::runtime_error(what) {}
};

class Setting {
 public:
  enum class Type { Group, List, Array, Int, Int64, Float, String, Bool };

  Setting() = default;
  explicit Setting(Type t) : type_(t) {}

  Type type() const { return type_; }
  const std::string &name() const { return name_; }
  bool is_group() const { return type_ == Type::Group; }
  bool is_list() const { return type_ == Type::List; }
  bool is_array() const { return type_ == Type::Array; }
  bool is_aggregate() const { return is_group() || is_list() || is_array(); }
  bool is_number() const { return type_ == Type::Int || type_ == Type::Int64 || type_ == Type::Float; }
  bool is_string() const { return type_ == Type::String; }

  // aggregates
  int length() const { return static_cast<int>(children_.size()); }
  bool exists(const std::string &path) const { return find(path) != nullptr; }
  const Setting &operator[](int i) const;
  const Setting &operator[](const std::string &name) const;   // throws ConfigError (SettingNotFoundException)
  const Setting &lookup(const std::string &dotted_path) const;
  const Setting *find(const std::string &dotted_path) const;  // nullptr if absent

  // scalars (AutoConvert between integer and float like libconfig's OptionAutoConvert)
  double as_double() const;
  long long as_int() const;
  bool as_bool() const;
  const std::string &as_string() const;

  // typed access with default, like jams::config_optional / config_required (interface/config.h:26-132)
  double get(const std::string &path, double dflt) const { const Setting *s = find(path); return s ? s->as_double() : dflt; }
  long long get(const std::string &path, long long dflt) const { const Setting *s = find(path); return s ? s->as_int() : dflt; }
  int get(const std::string &path, int dflt) const { const Setting *s = find(path); return s ? static_cast<int>(s->as_int()) : dflt; }
  bool get(const std::string &path, bool dflt) const { const Setting *s = find(path); return s ? s->as_bool() : dflt; }
  std::string get(const std::string &path, const std::string &dflt) const { const Setting *s = find(path); return s ? s->as_string() : dflt; }
  std::string get(const std::string &path, const char *dflt) const { return get(path, std::string(dflt)); }
  const Setting &required(const std::string &path) const;
  std::vector<double> doubles() const;   // array / list of numbers -> vector

  // construction (parser and merge)
  Setting &add(const std::string &name, Type t);
  Setting &add(Type t);
  void set_int(long long v, bool is64) { type_ = is64 ? Type::Int64 : Type::Int; i_ = v; }
  void set_float(double v) { type_ = Type::Float; d_ = v; }
  void set_string(const std::string &v) { type_ = Type::String; s_ = v; }
  void set_bool(bool v) { type_ = Type::Bool; b_ = v; }

  // JSON rendering (tests, provenance dumps)
  std::string to_json() const;

  // interface/config.cc:12-145
  static void overwrite(Setting &orig, const Setting &patch);

 private:
  Setting *child(const std::string &name);
  const Setting *child(const std::string &name) const;
  static void patch_aggregate(Setting &orig, const Setting &patch);
  static void patch_element(Setting &orig, const Setting &patch, int index);
  static void patch_simple(Setting &orig, const Setting &patch);
  void assign_scalar(const Setting &from);

  Type type_ = Type::Group;
  std::string name_;
  std::vector<std::unique_ptr<Setting>> children_;
  long long i_ = 0;
  double d_ = 0.0;
  bool b_ = false;
  std::string s_;
};

// parse one config text; throws ConfigError("line N: ...")
std::unique_ptr<Setting> parse_config_string(const std::string &text);
std::unique_ptr<Setting> parse_config_file(const std::string &filename);

// core/jams++.cc:48-84: each entry is a file name (if such a file exists) or a config string; merged left to right
std::unique_ptr<Setting> parse_config_strings(const std::vector<std::string> &args);

}  // namespace jams_b200

#endif  // JAMS_B200_HOST_CONFIG_H
