// tests/jams_stub/libconfig.h++ — TEST INFRASTRUCTURE: declarations of the part of libconfig++ (absent here) that the reference's headers
// and the JAMS-side adapter use, so that integration/jams/solvers/b200_llg_heun.cc can be COMPILED against the reference's real
// core/solver.h, hamiltonian/*.h and interface/config.h (tests/test_adapter_compile.py).  Nothing is defined: the object file is
// never linked.  Signatures follow libconfig++ 1.7's public header.
#ifndef JB_STUB_LIBCONFIG_HPP
#define JB_STUB_LIBCONFIG_HPP
#include <exception>
#include <string>
namespace libconfig {
class ConfigException : public std::exception {};
class SettingException : public ConfigException {
 public:
  const char *getPath() const;
};
class SettingTypeException : public SettingException {};
class SettingNotFoundException : public SettingException {};
class SettingNameException : public SettingException {};
class FileIOException : public ConfigException {};
class ParseException : public ConfigException {
 public:
  const char *getFile() const;
  int getLine() const;
  const char *getError() const;
};
class Setting {
 public:
  enum Type { TypeNone = 0, TypeInt, TypeInt64, TypeFloat, TypeString, TypeBoolean, TypeGroup, TypeArray, TypeList };
  enum Format { FormatDefault = 0, FormatHex = 1 };
  Type getType() const;
  operator bool() const;
  operator int() const;
  operator unsigned int() const;
  operator long() const;
  operator unsigned long() const;
  operator long long() const;
  operator unsigned long long() const;
  operator double() const;
  operator float() const;
  operator const char *() const;
  operator std::string() const;
  const char *c_str() const;
  Setting &operator=(bool);
  Setting &operator=(int);
  Setting &operator=(long);
  Setting &operator=(const long long &);
  Setting &operator=(const double &);
  Setting &operator=(float);
  Setting &operator=(const char *);
  Setting &operator=(const std::string &);
  Setting &lookup(const char *path) const;
  Setting &lookup(const std::string &path) const;
  Setting &operator[](const char *name) const;
  Setting &operator[](const std::string &name) const;
  Setting &operator[](int index) const;
  bool lookupValue(const char *name, bool &value) const;
  bool lookupValue(const char *name, int &value) const;
  bool lookupValue(const char *name, unsigned int &value) const;
  bool lookupValue(const char *name, long long &value) const;
  bool lookupValue(const char *name, double &value) const;
  bool lookupValue(const char *name, float &value) const;
  bool lookupValue(const char *name, const char *&value) const;
  bool lookupValue(const char *name, std::string &value) const;
  bool lookupValue(const std::string &name, bool &value) const;
  bool lookupValue(const std::string &name, int &value) const;
  bool lookupValue(const std::string &name, unsigned int &value) const;
  bool lookupValue(const std::string &name, long long &value) const;
  bool lookupValue(const std::string &name, double &value) const;
  bool lookupValue(const std::string &name, float &value) const;
  bool lookupValue(const std::string &name, const char *&value) const;
  bool lookupValue(const std::string &name, std::string &value) const;
  void remove(const char *name);
  void remove(const std::string &name);
  void remove(unsigned int idx);
  Setting &add(const char *name, Type type);
  Setting &add(const std::string &name, Type type);
  Setting &add(Type type);
  bool exists(const char *name) const;
  bool exists(const std::string &name) const;
  int getLength() const;
  const char *getName() const;
  std::string getPath() const;
  int getIndex() const;
  const Setting &getParent() const;
  Setting &getParent();
  bool isRoot() const;
  bool isGroup() const;
  bool isArray() const;
  bool isList() const;
  bool isAggregate() const;
  bool isScalar() const;
  bool isNumber() const;
  bool isString() const;
  unsigned int getSourceLine() const;
  const char *getSourceFile() const;
  Setting *begin();
  Setting *end();
  const Setting *begin() const;
  const Setting *end() const;
};
class Config {
 public:
  enum Option { OptionNone = 0, OptionAutoConvert = 0x01, OptionSemicolonSeparators = 0x02, OptionColonAssignmentForGroups = 0x04,
                OptionColonAssignmentForNonGroups = 0x08, OptionOpenBraceOnSeparateLine = 0x10, OptionAllowScientificNotation = 0x20,
                OptionFsync = 0x40, OptionAllowOverrides = 0x80 };
  Config();
  virtual ~Config();
  void setOptions(int options);
  int getOptions() const;
  void setOption(Config::Option option, bool flag);
  void readFile(const char *filename);
  void readFile(const std::string &filename);
  void writeFile(const char *filename);
  void readString(const char *str);
  void readString(const std::string &str);
  void write(FILE *stream) const;
  Setting &lookup(const std::string &path) const;
  Setting &lookup(const char *path) const;
  bool exists(const std::string &path) const;
  bool exists(const char *path) const;
  Setting &getRoot() const;
};
}  // namespace libconfig
#endif
