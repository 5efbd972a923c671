#include "ParameterReader.h"
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

static std::string trimmed(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}
static std::string lowered(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
  return s;
}
// the reference converts with a stringstream (arsenal stringToDouble): leading numeric prefix, else 0
static double to_double(const std::string& s) { return std::strtod(s.c_str(), nullptr); }

long ParameterReader::find(const std::string& name) const {
  const std::string key = lowered(trimmed(name));
  for (size_t i = 0; i < names.size(); i++) if (names[i] == key) return (long)i;
  return -1;
}

void ParameterReader::setVal(const std::string& name, double value) {
  const long idx = find(name);
  if (idx < 0) { names.push_back(lowered(trimmed(name))); values.push_back(value); }
  else values[idx] = value;
}

double ParameterReader::getVal(const std::string& name) const {
  const long idx = find(name);
  if (idx < 0) throw std::runtime_error("ParameterReader::getVal error: parameter with name " + name + " not found.");
  return values[idx];
}

void ParameterReader::phraseOneLine(const std::string& str, const std::string& commentSymbol) {
  if (trimmed(str).empty()) return;
  const std::string eq = str.substr(0, str.find(commentSymbol));
  if (trimmed(eq).empty()) return;
  const size_t pos = eq.find('=');
  if (pos == std::string::npos)
    throw std::runtime_error("ParameterReader: \"=\" symbol not found in equation assignment " + eq);
  setVal(eq.substr(0, pos), to_double(trimmed(eq.substr(pos + 1))));
}

void ParameterReader::readFromFile(const std::string& filename, const std::string& commentSymbol) {
  std::ifstream f(filename.c_str());
  if (!f) throw std::runtime_error("ParameterReader::readFromFile error: file " + filename + " does not exist.");
  std::string line;
  while (std::getline(f, line)) phraseOneLine(line, commentSymbol);
}

void ParameterReader::readFromArguments(long argc, char* argv[], const std::string& commentSymbol, long start_from) {
  for (long i = start_from; i < argc; i++) phraseOneLine(argv[i], commentSymbol);
}

void ParameterReader::echo() const {
  if (names.empty()) return;
  for (size_t i = 0; i < names.size(); i++) std::cout << names[i] << "=" << values[i] << "  ";
  std::cout << std::endl;
}
