// ParameterReader -- host-side mirror of the reference's configuration surface
// (reference src/ParameterReader.h:24-29, src/ParameterReader.cpp): `name = value  # comment` lines and
// `name=value` command-line overrides; names are trimmed and lower-cased, every value is a double,
// unknown names are accepted, a missing name is an error.  Unlike the reference (exit(-1)) a missing
// name throws, so the driver can report it and return a status.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

class ParameterReader {
 public:
  void readFromFile(const std::string& filename, const std::string& commentSymbol = "#");
  void readFromArguments(long argc, char* argv[], const std::string& commentSymbol = "#", long start_from = 1);
  void phraseOneLine(const std::string& str, const std::string& commentSymbol = "#");
  bool exist(const std::string& name) const { return find(name) >= 0; }
  void setVal(const std::string& name, double value);
  double getVal(const std::string& name) const;
  double getVal(const std::string& name, double fallback) const { return exist(name) ? getVal(name) : fallback; }
  void echo() const;
  size_t size() const { return names.size(); }

 private:
  long find(const std::string& name) const;
  std::vector<std::string> names;
  std::vector<double> values;
};
