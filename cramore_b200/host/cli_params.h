// cli_params.h -- `--flag value` parser with the surface of the reference's paramList
// (params.h:119-135, params.cpp:130-177, 449-486): long options only, value in the next argument
// (no `=`), repeated flags for multi-valued options, a second occurrence of a scalar flag is an
// error, unknown flags are collected and reported together, effective settings echoed to stderr.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace cli {

struct Option {
  enum Kind { GROUP, STRING, INT, DOUBLE, BOOL, MULTI_STRING, MULTI_DOUBLE } kind;
  std::string name, help;
  void* ptr;
  bool seen;
};

class Params {
 public:
  void group(const char* title) { opts_.push_back({Option::GROUP, title, "", nullptr, false}); }
  void add(const char* n, std::string* p, const char* h) { opts_.push_back({Option::STRING, n, h, p, false}); }
  void add(const char* n, int32_t* p, const char* h) { opts_.push_back({Option::INT, n, h, p, false}); }
  void add(const char* n, double* p, const char* h) { opts_.push_back({Option::DOUBLE, n, h, p, false}); }
  void add(const char* n, bool* p, const char* h) { opts_.push_back({Option::BOOL, n, h, p, false}); }
  void add(const char* n, std::vector<std::string>* p, const char* h) { opts_.push_back({Option::MULTI_STRING, n, h, p, false}); }
  void add(const char* n, std::vector<double>* p, const char* h) { opts_.push_back({Option::MULTI_DOUBLE, n, h, p, false}); }

  // returns false (after printing to stderr) on any error or on --help
  bool read(int argc, char** argv) {
    std::string errors;
    for (int i = 1; i < argc; ++i) {
      const char* a = argv[i];
      if (strcmp(a, "--help") == 0) {
        help();
        return false;
      }
      if (strncmp(a, "--", 2) != 0) {
        errors += std::string("Command line parameter ") + a + " (#" + std::to_string(i) + ") ignored\n";
        continue;
      }
      Option* o = find(a + 2);
      if (!o) {
        errors += std::string("Command line parameter ") + a + " (#" + std::to_string(i) + ") ignored\n";
        continue;
      }
      if (o->kind == Option::BOOL) {
        *(bool*)o->ptr = !*(bool*)o->ptr;
        o->seen = true;
        continue;
      }
      if (i + 1 >= argc) {
        errors += std::string("Missing value for --") + o->name + "\n";
        continue;
      }
      const char* v = argv[++i];
      const bool multi = o->kind == Option::MULTI_STRING || o->kind == Option::MULTI_DOUBLE;
      if (o->seen && !multi) {
        errors += std::string("Redundant use of option --") + o->name + " is not allowed\n";
        continue;
      }
      o->seen = true;
      char* end = nullptr;
      switch (o->kind) {
        case Option::STRING: *(std::string*)o->ptr = v; break;
        case Option::INT: {
          // params.cpp accepts decimal and 0x-prefixed values for int options (e.g. --excl-flag 0x0f04)
          long x = strtol(v, &end, 0);
          if (*end) errors += std::string("Invalid integer for --") + o->name + ": " + v + "\n";
          *(int32_t*)o->ptr = (int32_t)x;
          break;
        }
        case Option::DOUBLE: {
          double x = strtod(v, &end);
          if (*end) errors += std::string("Invalid number for --") + o->name + ": " + v + "\n";
          *(double*)o->ptr = x;
          break;
        }
        case Option::MULTI_STRING: ((std::vector<std::string>*)o->ptr)->push_back(v); break;
        case Option::MULTI_DOUBLE: {
          double x = strtod(v, &end);
          if (*end) errors += std::string("Invalid number for --") + o->name + ": " + v + "\n";
          ((std::vector<double>*)o->ptr)->push_back(x);
          break;
        }
        default: break;
      }
    }
    status();
    if (!errors.empty()) {
      fprintf(stderr, "\nFATAL ERROR - \nProblems encountered parsing command line:\n\n%s\n", errors.c_str());
      return false;
    }
    return true;
  }

  void status() const {
    fprintf(stderr, "\nThe following parameters are available.  Ones with \"[]\" are in effect:\n\n");
    for (const Option& o : opts_) {
      switch (o.kind) {
        case Option::GROUP: fprintf(stderr, "%s :\n", o.name.c_str()); break;
        case Option::STRING:
          if (((std::string*)o.ptr)->empty()) fprintf(stderr, "    --%s\n", o.name.c_str());
          else fprintf(stderr, "    --%s [%s]\n", o.name.c_str(), ((std::string*)o.ptr)->c_str());
          break;
        case Option::INT: fprintf(stderr, "    --%s [%d]\n", o.name.c_str(), *(int32_t*)o.ptr); break;
        case Option::DOUBLE: fprintf(stderr, "    --%s [%.2lf]\n", o.name.c_str(), *(double*)o.ptr); break;
        case Option::BOOL: fprintf(stderr, "    --%s [%s]\n", o.name.c_str(), *(bool*)o.ptr ? "ON" : "OFF"); break;
        case Option::MULTI_STRING: {
          fprintf(stderr, "    --%s", o.name.c_str());
          for (auto& s : *(std::vector<std::string>*)o.ptr) fprintf(stderr, " [%s]", s.c_str());
          fprintf(stderr, "\n");
          break;
        }
        case Option::MULTI_DOUBLE: {
          fprintf(stderr, "    --%s", o.name.c_str());
          for (double d : *(std::vector<double>*)o.ptr) fprintf(stderr, " [%.2lf]", d);
          fprintf(stderr, "\n");
          break;
        }
      }
    }
    fprintf(stderr, "\nRun with --help for more detailed help messages of each argument.\n\n");
  }

  void help() const {
    fprintf(stderr, "\nDetailed instructions of parameters are available. Ones with \"[]\" are in effect:\n\n");
    for (const Option& o : opts_) {
      if (o.kind == Option::GROUP) fprintf(stderr, "\n== %s ==\n", o.name.c_str());
      else fprintf(stderr, "   --%-20s : %s\n", o.name.c_str(), o.help.c_str());
    }
    fprintf(stderr, "\n");
  }

  // option names in declaration order (tests compare them with the reference's flag tables)
  std::vector<std::string> names() const {
    std::vector<std::string> v;
    for (const Option& o : opts_)
      if (o.kind != Option::GROUP) v.push_back(o.name);
    return v;
  }

 private:
  Option* find(const char* name) {
    for (Option& o : opts_)
      if (o.kind != Option::GROUP && o.name == name) return &o;
    return nullptr;
  }
  std::vector<Option> opts_;
};

}  // namespace cli
