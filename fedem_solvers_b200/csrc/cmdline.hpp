// cmdline.hpp -- command-line option handling with the semantics of the reference's FFaCmdLineArg
// (fedem-foundation/src/FFaLib/FFaCmdLineArg/FFaCmdLineArgImplementation.C:105-230,330-396 and the
// convertOption specialisations of FFaCmdLineArg.H:343-540), host code only:
//   * an argument starting with '-' is matched case-insensitively against the BEGINNING of every defined option
//     name, in alphabetical (std::map) order; "-opt=value" and "-optvalue" carry the value in the same argument
//     (the latter keeps searching for a longer option name, which is what makes -stressForm win over -stress);
//   * a bool option never consumes the next argument ("-opt" = "-opt+", "-opt-" switches off);
//   * other options take all following arguments up to the next one starting with '-' that is not a negative
//     number, joined by blanks;
//   * the first occurrence of an option wins: the command line is evaluated before the option files
//     (-fao, -fco, -fop, src/vpmCommon/cmdLineArgInitStd.C:103-135), whose content is appended;
//   * option files: blank-separated tokens, "quoted strings" kept together, '#' comments.
// Checked against the reference's own parser (oracle/_ref/libfedem_ref_ffl.so) in tests/test_cli_cpu.py.
#pragma once
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

namespace fsr {

class CmdLine {
 public:
  enum Type { BOOL, INT, DOUBLE, STRING };
  struct Option {
    Type type = STRING;
    bool b = false, bdef = false;
    int i = 0, idef = 0;
    double d = 0.0, ddef = 0.0;
    std::string s, sdef, help;
    bool is_set = false, is_public = true;
  };

  void add(const std::string& name, bool v, const char* help, bool pub = true) { Option o; o.type = BOOL; o.b = o.bdef = v; fin(name, o, help, pub); }
  void add(const std::string& name, int v, const char* help, bool pub = true) { Option o; o.type = INT; o.i = o.idef = v; fin(name, o, help, pub); }
  void add(const std::string& name, double v, const char* help, bool pub = true) { Option o; o.type = DOUBLE; o.d = o.ddef = v; fin(name, o, help, pub); }
  void add(const std::string& name, const char* v, const char* help, bool pub = true) { Option o; o.type = STRING; o.s = o.sdef = v; fin(name, o, help, pub); }

  void init(int argc, char** argv)
  {
    for (int i = 1; i < argc; ++i) args_.push_back(argv[i]);
    for (auto& o : opts_) { Option& p = o.second; p.b = p.bdef; p.i = p.idef; p.d = p.ddef; p.s = p.sdef; p.is_set = false; }
  }
  void push(const std::string& arg) { args_.push_back(arg); }
  void clear_options() { opts_.clear(); }   // keeps the (not yet evaluated) argument list

  bool read_options_file(const std::string& file)
  {
    if (file.empty()) return false;
    std::ifstream fs(file.c_str());
    if (!fs) { fprintf(stderr, " *** Could not open option file %s\n", file.c_str()); return false; }
    char c = ' ';
    std::string tmp;
    bool in_string = false;
    while (!fs.eof() && isspace((unsigned char)c)) fs.get(c);
    while (!fs.eof()) {
      if (c == '#' && !in_string) {
        fs.ignore(8192, '\n');
        if (tmp.empty()) { fs.get(c); continue; }
        c = ' ';
      }
      if (c == '"') in_string = !in_string;
      if (in_string || !isspace((unsigned char)c)) { tmp += c; fs.get(c); }
      else {
        args_.push_back(tmp);
        tmp.clear();
        while (!fs.eof() && isspace((unsigned char)c)) fs.get(c);
      }
    }
    if (!tmp.empty() && !in_string) args_.push_back(tmp);
    return true;
  }

  void evaluate()
  {
    for (size_t i = 0; i < args_.size(); ++i) {
      bool found = false;
      if (!args_[i].empty() && args_[i][0] == '-') {
        const std::string an = lower(args_[i].substr(1));
        for (auto& o : opts_) {
          const std::string key = lower(o.first);
          if (an.compare(0, key.size(), key) != 0 || an.size() < key.size()) continue;
          std::string argument;
          bool stop = found = true;
          const size_t n = o.first.size();
          if (an.size() > n) {
            if (an[n] == '=') argument = args_[i].substr(n + 2);
            else { argument = args_[i].substr(n + 1); stop = false; }
          } else if (o.second.type == BOOL)
            argument = "+";
          else
            while (i + 1 < args_.size() && (args_[i + 1].size() <= 1 || args_[i + 1][0] != '-' || isdigit((unsigned char)args_[i + 1][1])))
              argument = argument.empty() ? args_[++i] : argument + " " + args_[++i];
          if (o.second.is_set) continue;   // repeated option, the first instance counts
          const int invalid = convert(o.second, argument);
          if (invalid > 0) fprintf(stderr, "  ** Invalid option value for -%s: \"%s\" (ignored).\n", o.first.c_str(), argument.c_str());
          if (stop) break;
        }
      }
      if (!found) fprintf(stderr, "  ** Unknown command-line argument \"%s\" (ignored).\n", args_[i].c_str());
    }
    args_.clear();
  }

  const Option* find(const std::string& name) { evaluate(); auto it = opts_.find(name); return it == opts_.end() ? nullptr : &it->second; }
  bool get_bool(const std::string& n) { const Option* o = find(n); return o && o->type == BOOL ? o->b : false; }
  int get_int(const std::string& n) { const Option* o = find(n); return o && o->type == INT ? o->i : 0; }
  double get_double(const std::string& n) { const Option* o = find(n); return o && o->type == DOUBLE ? o->d : 0.0; }
  std::string get_string(const std::string& n) { const Option* o = find(n); return o && o->type == STRING ? o->s : std::string(); }
  bool is_set(const std::string& n) { const Option* o = find(n); return o ? o->is_set : false; }

  std::string help_text(bool all)
  {
    size_t longest = 0;
    for (auto& o : opts_) if (all || o.second.is_public) longest = std::max(longest, o.first.size());
    longest += 2;
    std::string t;
    for (auto& o : opts_) {
      if (!all && !o.second.is_public) continue;
      t += std::string(7, ' ') + "-" + o.first + std::string(longest - o.first.size(), ' ');
      for (char c : o.second.help) { t += c; if (c == '\n') t += std::string(longest + 8, ' '); }
      t += "\n" + std::string(longest + 8, ' ') + "Default: " + default_string(o.second) + "\n";
    }
    return t;
  }

 private:
  std::map<std::string, Option> opts_;
  std::vector<std::string> args_;

  void fin(const std::string& name, Option& o, const char* help, bool pub) { o.help = help ? help : ""; o.is_public = pub; opts_[name] = o; }
  static std::string lower(std::string s) { for (char& c : s) c = (char)tolower((unsigned char)c); return s; }
  static std::string default_string(const Option& o)
  {
    char b[64];
    switch (o.type) {
      case BOOL: return o.bdef ? "+ (true)" : "- (false)";
      case INT: snprintf(b, sizeof(b), "%d", o.idef); return b;
      case DOUBLE: snprintf(b, sizeof(b), "%g", o.ddef); return b;
      default: return o.sdef;
    }
  }
  // the convertOption specialisations: 0 = accepted or empty, > 0 = position of the offending character + 1
  static int convert(Option& o, const std::string& v)
  {
    char* e = nullptr;
    switch (o.type) {
      case BOOL:
        if (v.empty()) o.b = true;
        else if (v.size() != 1) return 0;   // (the reference returns `false` here: silently not set)
        else if (v == "+") o.b = true;
        else if (v == "-") o.b = false;
        else return 1;
        break;
      case INT: {
        if (v.empty()) return 0;
        const long a = strtol(v.c_str(), &e, 10);
        if (*e) return 1 + (int)(e - v.c_str());
        o.i = (int)a;
        break;
      }
      case DOUBLE: {
        if (v.empty()) return 0;
        const double a = strtod(v.c_str(), &e);
        if (*e) return 1 + (int)(e - v.c_str());
        o.d = a;
        break;
      }
      default:
        if (v.empty()) return 0;
        o.s = v;
        if (o.s.size() > 1) {
          if (o.s.front() == '"') o.s.erase(o.s.begin());
          if (o.s.back() == '"') o.s.erase(o.s.end() - 1);
        }
    }
    o.is_set = true;
    return 0;
  }
};

}  // namespace fsr
