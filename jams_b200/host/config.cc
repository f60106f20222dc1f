// config.cc — see config.h
#include "config.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace jams_b200 {

// ---------------------------------------------------------------------------------------------------
// Setting
// ---------------------------------------------------------------------------------------------------
Setting *Setting::child(const std::string &name) {
  for (auto &c : children_) if (c->name_ == name) return c.get();
  return nullptr;
}
const Setting *Setting::child(const std::string &name) const {
  for (auto &c : children_) if (c->name_ == name) return c.get();
  return nullptr;
}

const Setting *Setting::find(const std::string &path) const {
  const Setting *cur = this;
  size_t pos = 0;
  while (pos <= path.size()) {
    size_t dot = path.find('.', pos);
    if (dot == std::string::npos) dot = path.size();
    const std::string key = path.substr(pos, dot - pos);
    if (key.empty()) return nullptr;
    const Setting *next = nullptr;
    if (key[0] == '[' && key.back() == ']') {   // libconfig path syntax for elements: name.[index]
      const int idx = std::atoi(key.substr(1, key.size() - 2).c_str());
      if (idx >= 0 && idx < cur->length()) next = cur->children_[idx].get();
    } else if (cur->is_group()) {
      next = cur->child(key);
    }
    if (!next) return nullptr;
    cur = next;
    pos = dot + 1;
    if (dot == path.size()) break;
  }
  return cur;
}

const Setting &Setting::lookup(const std::string &path) const {
  const Setting *s = find(path);
  if (!s) throw ConfigError("setting not found: " + path);
  return *s;
}
const Setting &Setting::required(const std::string &path) const {
  const Setting *s = find(path);
  if (!s) throw ConfigError("required setting '" + path + "' is missing" + (name_.empty() ? "" : " in '" + name_ + "'"));
  return *s;
}
const Setting &Setting::operator[](const std::string &name) const { return lookup(name); }
const Setting &Setting::operator[](int i) const {
  if (i < 0 || i >= length()) throw ConfigError("setting index " + std::to_string(i) + " out of range" + (name_.empty() ? "" : " in '" + name_ + "'"));
  return *children_[i];
}

double Setting::as_double() const {
  switch (type_) {
    case Type::Float: return d_;
    case Type::Int: case Type::Int64: return static_cast<double>(i_);
    default: throw ConfigError("setting '" + name_ + "' is not a number");
  }
}
long long Setting::as_int() const {
  switch (type_) {
    case Type::Int: case Type::Int64: return i_;
    case Type::Float: return static_cast<long long>(d_);   // AutoConvert truncates like a C cast
    default: throw ConfigError("setting '" + name_ + "' is not a number");
  }
}
bool Setting::as_bool() const {
  if (type_ != Type::Bool) throw ConfigError("setting '" + name_ + "' is not a boolean");
  return b_;
}
const std::string &Setting::as_string() const {
  if (type_ != Type::String) throw ConfigError("setting '" + name_ + "' is not a string");
  return s_;
}
std::vector<double> Setting::doubles() const {
  if (!is_array() && !is_list()) throw ConfigError("setting '" + name_ + "' is not an array");
  std::vector<double> v;
  for (auto &c : children_) v.push_back(c->as_double());
  return v;
}

Setting &Setting::add(const std::string &name, Type t) {
  if (!is_group()) throw ConfigError("cannot add a named setting to a non-group");
  if (child(name)) throw ConfigError("duplicate setting name: " + name);
  children_.emplace_back(new Setting(t));
  children_.back()->name_ = name;
  return *children_.back();
}
Setting &Setting::add(Type t) {
  if (!is_list() && !is_array()) throw ConfigError("cannot add an unnamed setting to a group");
  children_.emplace_back(new Setting(t));
  return *children_.back();
}

static void json_escape(const std::string &s, std::string &out) {
  out += '"';
  for (char ch : s) {
    switch (ch) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\t': out += "\\t"; break;
      case '\r': out += "\\r"; break;
      default:
        if (static_cast<unsigned char>(ch) < 0x20) { char buf[8]; std::snprintf(buf, sizeof buf, "\\u%04x", ch); out += buf; }
        else out += ch;
    }
  }
  out += '"';
}

std::string Setting::to_json() const {
  std::string out;
  switch (type_) {
    case Type::Group:
      out += '{';
      for (size_t i = 0; i < children_.size(); ++i) {
        if (i) out += ", ";
        json_escape(children_[i]->name_, out);
        out += ": " + children_[i]->to_json();
      }
      out += '}';
      break;
    case Type::List: case Type::Array:
      out += '[';
      for (size_t i = 0; i < children_.size(); ++i) { if (i) out += ", "; out += children_[i]->to_json(); }
      out += ']';
      break;
    case Type::Int: case Type::Int64: out += std::to_string(i_); break;
    case Type::Float: {
      char buf[40];
      std::snprintf(buf, sizeof buf, "%.17g", d_);
      out += buf;
      if (out.find_first_of(".eEn") == std::string::npos) out += ".0";
      break;
    }
    case Type::String: json_escape(s_, out); break;
    case Type::Bool: out += b_ ? "true" : "false"; break;
  }
  return out;
}

// ---- merge (interface/config.cc:12-145) -------------------------------------------------------------
void Setting::assign_scalar(const Setting &from) {
  type_ = from.type_; i_ = from.i_; d_ = from.d_; b_ = from.b_; s_ = from.s_;
  children_.clear();
}

// config_patch_simple (:15-52): `patch` is a member of a group
void Setting::patch_simple(Setting &orig, const Setting &patch) {
  if (patch.is_aggregate()) {
    // config_patch_add_or_merge_aggregate (:54-83), group branch
    Setting *agg = orig.is_group() ? orig.child(patch.name_) : nullptr;
    if (!agg) agg = &orig.add(patch.name_, patch.type_);
    patch_aggregate(*agg, patch);
    return;
  }
  if (Setting *old = orig.child(patch.name_)) {   // orig.remove(name); orig.add(name, type) = value
    for (auto it = orig.children_.begin(); it != orig.children_.end(); ++it) if (it->get() == old) { orig.children_.erase(it); break; }
  }
  orig.add(patch.name_, patch.type_).assign_scalar(patch);
}

// config_patch_element (:100-143): `patch` is element `index` of a list or array
void Setting::patch_element(Setting &orig, const Setting &patch, int index) {
  if (patch.is_aggregate()) {
    while (orig.length() <= index) orig.add(patch.type_);   // (:67-76) extend so that the index exists
    patch_aggregate(*orig.children_[index], patch);
    return;
  }
  Setting *target = index < orig.length() ? orig.children_[index].get() : &orig.add(patch.type_);
  target->assign_scalar(patch);
}

// config_patch_aggregate (:87-98) once the target aggregate has been found
void Setting::patch_aggregate(Setting &aggregate, const Setting &patch) {
  for (int i = 0; i < patch.length(); ++i) {
    if (patch.is_group()) {
      if (!aggregate.is_group()) throw ConfigError("config patch: '" + patch.name_ + "' is a group but the original setting is not");
      patch_simple(aggregate, *patch.children_[i]);
    } else {
      if (aggregate.is_group()) throw ConfigError("config patch: '" + patch.name_ + "' is a list but the original setting is a group");
      patch_element(aggregate, *patch.children_[i], i);
    }
  }
}

void Setting::overwrite(Setting &orig, const Setting &patch) {
  if (!orig.is_group() && !orig.is_list()) return;   // (:146-148)
  patch_aggregate(orig, patch);
}

// ---------------------------------------------------------------------------------------------------
// parser
// ---------------------------------------------------------------------------------------------------
namespace {

class Parser {
 public:
  explicit Parser(const std::string &t) : t_(t) {}

  std::unique_ptr<Setting> parse() {
    std::unique_ptr<Setting> root(new Setting(Setting::Type::Group));
    parse_settings(*root, /*closing=*/'\0');
    return root;
  }

 private:
  [[noreturn]] void fail(const std::string &msg) const { throw ConfigError("line " + std::to_string(line_) + ": " + msg); }

  void skip_ws() {
    for (;;) {
      while (p_ < t_.size() && std::isspace(static_cast<unsigned char>(t_[p_]))) { if (t_[p_] == '\n') ++line_; ++p_; }
      if (p_ >= t_.size()) return;
      if (t_[p_] == '#' || (t_[p_] == '/' && p_ + 1 < t_.size() && t_[p_ + 1] == '/')) {
        while (p_ < t_.size() && t_[p_] != '\n') ++p_;
      } else if (t_[p_] == '/' && p_ + 1 < t_.size() && t_[p_ + 1] == '*') {
        p_ += 2;
        while (p_ + 1 < t_.size() && !(t_[p_] == '*' && t_[p_ + 1] == '/')) { if (t_[p_] == '\n') ++line_; ++p_; }
        if (p_ + 1 >= t_.size()) fail("unterminated comment");
        p_ += 2;
      } else {
        return;
      }
    }
  }

  bool at_end() { skip_ws(); return p_ >= t_.size(); }
  char peek() { skip_ws(); return p_ < t_.size() ? t_[p_] : '\0'; }

  std::string parse_name() {
    skip_ws();
    size_t b = p_;
    if (p_ < t_.size() && (std::isalpha(static_cast<unsigned char>(t_[p_])) || t_[p_] == '*' || t_[p_] == '_')) {
      ++p_;
      while (p_ < t_.size() && (std::isalnum(static_cast<unsigned char>(t_[p_])) || t_[p_] == '_' || t_[p_] == '-' || t_[p_] == '*')) ++p_;
    }
    if (b == p_) fail("syntax error: setting name expected");
    return t_.substr(b, p_ - b);
  }

  void parse_settings(Setting &group, char closing) {
    for (;;) {
      if (at_end()) { if (closing) fail("unexpected end of input: missing '}'"); return; }
      if (closing && peek() == closing) { ++p_; return; }
      const std::string name = parse_name();
      const char c = peek();
      if (c != '=' && c != ':') fail("syntax error: '=' or ':' expected after '" + name + "'");
      ++p_;
      if (group.exists(name)) fail("duplicate setting name '" + name + "'");
      parse_value(group, &name);
      const char e = peek();
      if (e == ';' || e == ',') ++p_;
    }
  }

  // parses one value and adds it to `parent` (named if name != nullptr)
  void parse_value(Setting &parent, const std::string *name) {
    const char c = peek();
    auto make = [&](Setting::Type t) -> Setting & { return name ? parent.add(*name, t) : parent.add(t); };
    if (c == '{') {
      ++p_;
      parse_settings(make(Setting::Type::Group), '}');
    } else if (c == '(') {
      ++p_;
      Setting &list = make(Setting::Type::List);
      parse_elements(list, ')');
    } else if (c == '[') {
      ++p_;
      Setting &arr = make(Setting::Type::Array);
      parse_elements(arr, ']');
      for (int i = 1; i < arr.length(); ++i) {   // arrays are homogeneous scalars; AutoConvert lets ints and floats mix
        const bool num0 = arr[0].is_number(), numi = arr[i].is_number();
        if (arr[i].is_aggregate() || (num0 != numi) || (!num0 && arr[i].type() != arr[0].type())) fail("mismatched element type in array");
      }
    } else if (c == '"') {
      std::string s = parse_string();
      while (peek() == '"') s += parse_string();   // adjacent string literals concatenate
      make(Setting::Type::String).set_string(s);
    } else {
      parse_scalar(make(Setting::Type::Int));
    }
  }

  void parse_elements(Setting &agg, char closing) {
    for (;;) {
      if (at_end()) fail(std::string("unexpected end of input: missing '") + closing + "'");
      if (peek() == closing) { ++p_; return; }
      parse_value(agg, nullptr);
      const char e = peek();
      if (e == ',') { ++p_; continue; }
      if (e == closing) { ++p_; return; }
      fail(std::string("syntax error: ',' or '") + closing + "' expected");
    }
  }

  std::string parse_string() {
    std::string s;
    ++p_;   // opening quote
    for (;;) {
      if (p_ >= t_.size()) fail("unterminated string");
      char ch = t_[p_++];
      if (ch == '"') break;
      if (ch == '\n') ++line_;
      if (ch == '\\') {
        if (p_ >= t_.size()) fail("unterminated string");
        const char e = t_[p_++];
        switch (e) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'f': s += '\f'; break;
          case '\\': s += '\\'; break;
          case '"': s += '"'; break;
          case 'x': {
            if (p_ + 1 >= t_.size()) fail("bad \\x escape");
            s += static_cast<char>(std::strtol(t_.substr(p_, 2).c_str(), nullptr, 16));
            p_ += 2;
            break;
          }
          default: fail(std::string("unknown escape \\") + e);
        }
      } else {
        s += ch;
      }
    }
    return s;
  }

  void parse_scalar(Setting &s) {
    skip_ws();
    size_t b = p_;
    while (p_ < t_.size() && (std::isalnum(static_cast<unsigned char>(t_[p_])) || t_[p_] == '+' || t_[p_] == '-' || t_[p_] == '.' || t_[p_] == '_')) ++p_;
    std::string tok = t_.substr(b, p_ - b);
    if (tok.empty()) fail(std::string("syntax error near '") + (p_ < t_.size() ? t_[p_] : ' ') + "'");
    std::string low;
    for (char ch : tok) low += static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    if (low == "true") { s.set_bool(true); return; }
    if (low == "false") { s.set_bool(false); return; }
    // integer: [-+]?digits (L|LL)? or 0x hex
    bool is64 = false;
    std::string body = tok;
    while (!body.empty() && (body.back() == 'L' || body.back() == 'l')) { body.pop_back(); is64 = true; }
    char *end = nullptr;
    if (body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'X')) {
      const unsigned long long v = std::strtoull(body.c_str(), &end, 16);
      if (*end == '\0') { s.set_int(static_cast<long long>(v), is64 || v > 0x7fffffffULL); return; }
    }
    const bool looks_int = body.find_first_of(".eE") == std::string::npos;
    if (looks_int) {
      const long long v = std::strtoll(body.c_str(), &end, 10);
      if (end != body.c_str() && *end == '\0') { s.set_int(v, is64 || v > 2147483647LL || v < -2147483648LL); return; }
    } else if (!is64) {
      const double v = std::strtod(tok.c_str(), &end);
      if (end != tok.c_str() && *end == '\0') { s.set_float(v); return; }
    }
    fail("syntax error: cannot parse value '" + tok + "'");
  }

  const std::string &t_;
  size_t p_ = 0;
  int line_ = 1;
};

}  // namespace

std::unique_ptr<Setting> parse_config_string(const std::string &text) { return Parser(text).parse(); }

std::unique_ptr<Setting> parse_config_file(const std::string &filename) {
  std::ifstream f(filename);
  if (!f) throw ConfigError("IO error opening config file: " + filename);
  std::stringstream ss;
  ss << f.rdbuf();
  try {
    return parse_config_string(ss.str());
  } catch (const ConfigError &e) {
    throw ConfigError("Error parsing config file: " + filename + ":" + e.what());
  }
}

std::unique_ptr<Setting> parse_config_strings(const std::vector<std::string> &args) {
  std::unique_ptr<Setting> combined(new Setting(Setting::Type::Group));
  for (const auto &s : args) {
    std::unique_ptr<Setting> patch;
    std::ifstream probe(s);
    if (probe.good()) {
      probe.close();
      patch = parse_config_file(s);
    } else {
      try {
        patch = parse_config_string(s);
      } catch (const ConfigError &e) {
        throw ConfigError("File not found or error parsing config string:\n  '" + s + "'\n" + e.what());
      }
    }
    Setting::overwrite(*combined, *patch);
  }
  return combined;
}

}  // namespace jams_b200
