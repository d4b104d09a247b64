// MPS (free and fixed format) -> lp_problem_t.
//
// Independent implementation of the ingest step the reference performs with
// cpp/libmps_parser (parse_mps<int,double>, called from cuOptReadProblem,
// cpp/src/linear_programming/cuopt_c.cpp:62-86).  Behaviour is matched to
// libmps_parser/src/mps_parser.cpp on every file under the reference's
// datasets/ (tests/test_mps_reader.py compares against golden dumps produced by
// the reference parser itself):
//   * sections NAME, OBJSENSE, OBJNAME, ROWS (+LAZYCONS), COLUMNS (with
//     'MARKER' INTORG/INTEND), RHS, RANGES, BOUNDS, ENDATA      (mps_parser.cpp:318-447)
//   * first N row = objective, further N rows ignored            (:519-538)
//   * RHS on the objective row = minus the objective offset       (:718-722)
//   * RANGES semantics per row type                               (:183-236)
//   * bound types LO UP FX FR MI PL BV LI UI; negative UP with no
//     earlier bound opens the lower bound; unbounded integers -> [0,1]   (:969-1051, :480-494)
//   * the set-name field of RHS / BOUNDS is optional in free format (:676-689, :760-776)
//
// Design: each data line is cut into fields once (whitespace tokens in free
// format, column slices in fixed format) and every section consumes the field
// list, instead of the reference's positional re-scanning.
#include "lp_problem.hpp"

#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace cuopt_b200 {
namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

[[noreturn]] void parse_fail(const std::string& msg) { throw lp_error(error_type_t::MpsParseError, msg); }

std::string_view trim(std::string_view s)
{
  const auto b = s.find_first_not_of(" \t\r");
  if (b == std::string_view::npos) return {};
  const auto e = s.find_last_not_of(" \t\r");
  return s.substr(b, e - b + 1);
}

std::string_view slice(std::string_view s, size_t pos, size_t len)
{
  if (pos >= s.size()) return {};
  return s.substr(pos, len);
}

double to_number(std::string_view tok, const char* where, std::string_view line)
{
  // std::stod semantics (longest valid prefix, "inf"/"nan" accepted), as the reference uses it.
  try {
    return std::stod(std::string(tok));
  } catch (const std::exception&) {
    parse_fail(std::string("Bad value found in ") + where + "! line=" + std::string(line));
  }
}

enum class section_t { None, Rows, Columns, Rhs, Bounds, Ranges, ObjSense, ObjName };

struct reader_t {
  bool fixed;
  lp_problem_t out;

  std::vector<char> row_kind;  // 'E','L','G' (or whatever the file said; validated at the end)
  std::vector<std::vector<int>> row_cols;
  std::vector<std::vector<double>> row_vals;
  std::vector<double> rhs, ranges;
  std::unordered_map<std::string, int> row_id, var_id;
  std::unordered_set<std::string> ignored_objectives, seen;
  std::unordered_set<int> var_has_bound;
  bool in_integer_block = false;

  explicit reader_t(bool fixed_format) : fixed(fixed_format) {}

  // ---- field splitting ----------------------------------------------------
  // Free format: whitespace-separated tokens.  Fixed format: the classic card
  // layout [2-3] [5-12] [15-22] [25-36] [40-47] [50-61]; empty trailing fields dropped.
  std::vector<std::string_view> fields(std::string_view line) const
  {
    std::vector<std::string_view> f;
    if (!fixed) {
      size_t p = 0;
      while (true) {
        p = line.find_first_not_of(" \t\r", p);
        if (p == std::string_view::npos) break;
        size_t e = line.find_first_of(" \t\r", p);
        if (e == std::string_view::npos) e = line.size();
        f.push_back(line.substr(p, e - p));
        p = e;
      }
    } else {
      static const size_t pos[6] = {1, 4, 14, 24, 39, 49};
      static const size_t len[6] = {2, 8, 8, 12, 8, 12};
      for (int i = 0; i < 6; ++i)
        f.push_back(trim(slice(line, pos[i], len[i])));
      // the second (name, value) pair counts only when something lies beyond column 40, exactly the
      // reference's test `line.find_last_not_of(" \r\t\n") > 39` (mps_parser.cpp:603, :692, :881): a first
      // value wider than its 12-character field must not be mistaken for a second pair
      const auto last = line.find_last_not_of(" \t\r\n");
      if (last == std::string_view::npos || last <= 39) f.resize(4);
      while (!f.empty() && f.back().empty())
        f.pop_back();
    }
    return f;
  }

  // ---- sections -------------------------------------------------------------
  void on_row(std::string_view line)
  {
    char kind;
    std::string name;
    if (fixed) {
      kind = line.size() > 1 ? line[1] : ' ';
      name = std::string(trim(slice(line, 4, 8)));
    } else {
      auto f = fields(line);
      kind   = f.empty() ? ' ' : f[0][0];
      name   = f.size() > 1 ? std::string(f[1]) : std::string();
    }
    if (kind == 'N') {
      if (out.objective_name.empty())
        out.objective_name = name;
      else
        ignored_objectives.insert(name);
      return;
    }
    if (row_id.count(name)) parse_fail("Duplicate row named '" + name + "' found! line=" + std::string(line));
    row_id.emplace(name, (int)out.row_names.size());
    out.row_names.push_back(name);
    row_kind.push_back(kind);
  }

  void begin_columns()
  {
    row_cols.resize(out.row_names.size());
    row_vals.resize(out.row_names.size());
    rhs.assign(out.row_names.size(), 0.0);
  }

  void add_entry(std::string_view line, int var, std::string_view row, std::string_view num)
  {
    if (ignored_objectives.count(std::string(row))) return;
    const double v = to_number(num, "COLUMNS", line);
    if (row == out.objective_name) {
      out.objective_coefficients[var] = v;
      return;
    }
    auto it = row_id.find(std::string(row));
    if (it == row_id.end()) parse_fail("Bad row name found '" + std::string(row) + "' in COLUMNS! line=" + std::string(line));
    row_cols[it->second].push_back(var);
    row_vals[it->second].push_back(v);
  }

  void on_column(std::string_view line)
  {
    if (fixed && line.size() < 25) parse_fail("COLUMNS should have atleast 3 entities! line=" + std::string(line));
    std::string_view name;
    std::vector<std::string_view> rest;  // (row, value) pairs
    if (fixed) {
      auto f = fields(line);
      name   = f.size() > 1 ? f[1] : std::string_view{};
      for (size_t i = 2; i < f.size(); ++i)
        rest.push_back(f[i]);
    } else {
      auto f = fields(line);
      if (f.empty()) return;
      name = f[0];
      rest.assign(f.begin() + 1, f.end());
    }
    if (line.find("'MARKER'") != std::string_view::npos) {
      if (line.find("INTORG") != std::string_view::npos) {
        if (in_integer_block) parse_fail("Cannot capture an int section while already capturing an int section");
        in_integer_block = true;
      }
      if (line.find("INTEND") != std::string_view::npos) {
        if (!in_integer_block) parse_fail("Cannot stop int capture when a previous capture is not started");
        in_integer_block = false;
      }
      return;
    }
    if (out.variable_names.empty() || out.variable_names.back() != name) {
      if (var_id.count(std::string(name)))
        parse_fail("All rows for the column (" + std::string(name) + ") should occur contiguously! line=" + std::string(line));
      var_id.emplace(std::string(name), (int)out.variable_names.size());
      out.variable_names.emplace_back(name);
      out.variable_types.push_back(in_integer_block ? 'I' : 'C');
      out.objective_coefficients.push_back(0.0);
    }
    const int var = (int)out.variable_names.size() - 1;
    for (size_t i = 0; i < 2 && 2 * i < rest.size(); ++i) {
      std::string_view row = rest[2 * i];
      if (row.empty() || row[0] == '$') return;
      std::string_view num = 2 * i + 1 < rest.size() ? rest[2 * i + 1] : std::string_view{};
      add_entry(line, var, row, num);
    }
  }

  void set_rhs(std::string_view line, std::string_view row, std::string_view num)
  {
    const double v = to_number(num, "RHS", line);
    if (row == out.objective_name) {
      out.objective_offset = -v;  // RHS of the objective row is minus the constant term
      return;
    }
    auto it = row_id.find(std::string(row));
    if (it == row_id.end()) parse_fail("Bad row name found '" + std::string(row) + "' in RHS! line=" + std::string(line));
    rhs[it->second] = v;
  }

  void on_rhs(std::string_view line)
  {
    if (fixed && line.size() < 25) parse_fail("RHS should have atleast 3 entities! line=" + std::string(line));
    auto f = fields(line);
    size_t start;
    if (fixed) {
      start = 2;  // field 2 is the (ignored) RHS set name
    } else {
      if (f.empty()) return;
      const bool first_is_row = f[0] == out.objective_name || row_id.count(std::string(f[0]));
      start                   = first_is_row ? 0 : 1;
    }
    for (size_t i = 0; i < 2 && start + 2 * i < f.size(); ++i) {
      std::string_view row = f[start + 2 * i];
      if (row.empty() || row[0] == '$') return;
      std::string_view num = start + 2 * i + 1 < f.size() ? f[start + 2 * i + 1] : std::string_view{};
      set_rhs(line, row, num);
    }
  }

  void begin_bounds()
  {
    out.variable_lower_bounds.assign(out.variable_names.size(), 0.0);
    out.variable_upper_bounds.assign(out.variable_names.size(), kInf);
  }

  void on_bound(std::string_view line)
  {
    if (fixed && line.size() < 14) parse_fail("BOUNDS should have atleast 2 entities! line=" + std::string(line));
    auto f = fields(line);
    if (f.empty()) return;
    const std::string_view type = f[0];
    enum { LO, UP, FX, FR, MI, PL, BV, LI, UI } kind;
    if (type == "LO") kind = LO;
    else if (type == "UP") kind = UP;
    else if (type == "FX") kind = FX;
    else if (type == "FR") kind = FR;
    else if (type == "MI") kind = MI;
    else if (type == "PL") kind = PL;
    else if (type == "BV") kind = BV;
    else if (type == "LI") kind = LI;
    else if (type == "UI") kind = UI;
    else if (type == "LC") parse_fail("Unsupported semi continous bound type found! Line=" + std::string(line));
    else parse_fail("Invalid variable bound type found in BOUNDS section! Bound type=" + std::string(type));

    std::string_view var, num;
    if (fixed) {
      var = f.size() > 2 ? f[2] : std::string_view{};
      num = f.size() > 3 ? f[3] : std::string_view{};
    } else {
      // "TYPE [set] var [value]": the set name is absent when field 1 is already a variable.
      size_t vi = (f.size() > 1 && var_id.count(std::string(f[1]))) ? 1 : 2;
      var       = f.size() > vi ? f[vi] : std::string_view{};
      num       = f.size() > vi + 1 ? f[vi + 1] : std::string_view{};
    }
    if (!var.empty() && var[0] == '$') return;

    auto it = var_id.find(std::string(var));
    if (it == var_id.end()) {
      // A variable first mentioned in BOUNDS: a column with no entries (mps_parser.cpp:783-793).
      it = var_id.emplace(std::string(var), (int)out.variable_names.size()).first;
      out.variable_names.emplace_back(var);
      out.objective_coefficients.push_back(0.0);
      out.variable_lower_bounds.push_back(0.0);
      out.variable_upper_bounds.push_back(kInf);
      out.variable_types.push_back('C');
    }
    const int j       = it->second;
    double& lo        = out.variable_lower_bounds[j];
    double& hi        = out.variable_upper_bounds[j];
    const bool first  = !var_has_bound.count(j);
    auto value        = [&]() { return to_number(num, "BOUNDS", line); };
    switch (kind) {
      case LO: lo = value(); break;
      case UP:
        hi = value();
        if (first && hi < 0.0) lo = -kInf;
        break;
      case FX: lo = hi = value(); break;
      case FR: lo = -kInf; hi = kInf; break;
      case MI: lo = -kInf; break;
      case PL: hi = kInf; break;
      case BV: lo = 0.0; hi = 1.0; out.variable_types[j] = 'I'; break;
      case LI:
        if (first) hi = kInf;
        lo                    = value();
        out.variable_types[j] = 'I';
        break;
      case UI:
        hi = value();
        if (first && hi < 0.0) lo = -kInf;
        out.variable_types[j] = 'I';
        break;
    }
    var_has_bound.insert(j);
  }

  void begin_ranges() { ranges.assign(row_kind.size(), kInf); }  // +inf == "no range given"

  void set_range(std::string_view line, std::string_view row, std::string_view num)
  {
    const double v = to_number(num, "RANGES", line);
    auto it        = row_id.find(std::string(row));
    if (it == row_id.end()) parse_fail("Bad row name found '" + std::string(row) + "' in RANGES! line=" + std::string(line));
    ranges[it->second] = v;
  }

  void on_range(std::string_view line)
  {
    if (fixed && line.size() < 25) parse_fail("RANGES should have atleast 2 entities! line=" + std::string(line));
    auto f             = fields(line);
    const size_t start = fixed ? 2 : 1;  // field before it is the (mandatory) RANGES set name
    for (size_t i = 0; i < 2 && start + 2 * i < f.size(); ++i) {
      std::string_view row = f[start + 2 * i];
      if (row.empty() || row[0] == '$') return;
      std::string_view num = start + 2 * i + 1 < f.size() ? f[start + 2 * i + 1] : std::string_view{};
      set_range(line, row, num);
    }
  }

  void on_objsense(std::string_view line)
  {
    if (fixed) parse_fail("OBJSENSE only exist in Free MPS format");
    auto f = reader_t(false).fields(line);
    size_t i = (!f.empty() && f[0] == "OBJSENSE") ? 1 : 0;
    std::string_view w = i < f.size() ? f[i] : std::string_view{};
    if (w == "MIN" || w == "MINIMIZE") out.maximize = false;
    else if (w == "MAX" || w == "MAXIMIZE") out.maximize = true;
    else parse_fail("Invalid variable bound type found in OBJSENSE section! Objsense type=" + std::string(w));
  }

  void on_objname(std::string_view line)
  {
    if (fixed) parse_fail("OBJNAME only exist in Free MPS format");
    auto f = reader_t(false).fields(line);
    size_t i = (!f.empty() && f[0] == "OBJNAME") ? 1 : 0;
    if (!out.objective_name.empty()) parse_fail("OBJNAME section should appear before ROWS section");
    out.objective_name = i < f.size() ? std::string(f[i]) : std::string();
  }

  static bool has_alpha(std::string_view s)
  {
    for (char c : s)
      if (std::isalpha((unsigned char)c)) return true;
    return false;
  }

  void run(std::string_view text)
  {
    if (text.find('\n') == std::string_view::npos && text.empty())
      parse_fail("Error parsing MPS file! No line return found (\"\\n\")");
    section_t sec = section_t::None;
    size_t p      = 0;
    bool any_line = false;
    while (p < text.size()) {
      size_t e = text.find('\n', p);
      if (e == std::string_view::npos) e = text.size();
      std::string_view line = text.substr(p, e - p);
      p                     = e + 1;
      if (line.empty()) continue;
      any_line = true;
      if (line[0] == '*' || line[0] == '$' || line[0] == '\r') continue;
      if (line[0] != ' ' && line[0] != '\t') {
        auto starts = [&](const char* kw) { return line.compare(0, std::strlen(kw), kw) == 0; };
        if (starts("NAME")) {
          seen.insert("NAME");
          const auto b = line.find_first_not_of(" \t", 4);
          if (b != std::string_view::npos) {
            if (fixed) {
              out.problem_name = std::string(trim(line.substr(b, 8)));
            } else {
              auto f           = fields(line.substr(b));
              out.problem_name = f.empty() ? std::string() : std::string(f[0]);
            }
          }
        } else if (starts("ROWS")) {
          seen.insert("ROWS");
          sec = section_t::Rows;
        } else if (starts("COLUMNS")) {
          seen.insert("COLUMNS");
          sec = section_t::Columns;
          begin_columns();
        } else if (starts("RHS")) {
          seen.insert("RHS");
          sec = section_t::Rhs;
        } else if (starts("BOUNDS")) {
          seen.insert("BOUNDS");
          sec = section_t::Bounds;
          begin_bounds();
        } else if (starts("RANGES")) {
          seen.insert("RANGES");
          sec = section_t::Ranges;
          begin_ranges();
        } else if (starts("OBJSENSE")) {
          if (has_alpha(line.substr(8))) {
            on_objsense(line);  // direction on the header line itself
          } else {
            seen.insert("OBJSENSE");
            sec = section_t::ObjSense;
          }
        } else if (starts("OBJNAME")) {
          seen.insert("OBJNAME");
          if (has_alpha(line.substr(7))) {
            on_objname(line);
          } else {
            sec = section_t::ObjName;
          }
        } else if (starts("ENDATA")) {
          seen.insert("ENDATA");
          break;
        } else if (starts("LAZYCONS")) {
          seen.insert("LAZYCONS");
          sec = section_t::Rows;  // lazy constraints are ordinary rows
        } else {
          parse_fail("Invalid named block found! Line=" + std::string(line));
        }
        continue;
      }
      switch (sec) {
        case section_t::Rows: on_row(line); break;
        case section_t::Columns: on_column(line); break;
        case section_t::Rhs: on_rhs(line); break;
        case section_t::Bounds: on_bound(line); break;
        case section_t::Ranges: on_range(line); break;
        case section_t::ObjSense: on_objsense(line); break;
        case section_t::ObjName: on_objname(line); break;
        default: parse_fail("Ended up at a bad parser state! Line=" + std::string(line));
      }
    }
    if (!any_line) parse_fail("Error parsing MPS file! No line return found (\"\\n\")");
    finish();
  }

  void finish()
  {
    if (out.objective_name.empty()) parse_fail("No objective found!");
    if (!seen.count("ROWS")) parse_fail("ROWS section is missing");
    if (!seen.count("COLUMNS")) parse_fail("COLUMNS section is missing");
    if (!seen.count("RHS")) parse_fail("RHS section is missing");

    const size_t n = out.variable_names.size();
    if (out.variable_upper_bounds.empty()) begin_bounds();
    if (out.variable_lower_bounds.size() != n || out.variable_upper_bounds.size() != n)
      parse_fail("MPS reader internal error: bound vector sizes");
    for (size_t j = 0; j < n; ++j) {
      if (!var_has_bound.count((int)j) && out.variable_types[j] == 'I') {
        out.variable_lower_bounds[j] = 0.0;
        out.variable_upper_bounds[j] = 1.0;
      }
      if (!(out.variable_lower_bounds[j] <= out.variable_upper_bounds[j]))
        parse_fail("Variable " + out.variable_names[j] + " has lower bound above upper bound");
    }

    const size_t m    = out.row_names.size();
    out.n_constraints = (int)m;
    out.n_variables   = (int)n;
    row_cols.resize(m);
    row_vals.resize(m);
    rhs.resize(m, 0.0);
    out.A_offsets.assign(1, 0);
    for (size_t i = 0; i < m; ++i) {
      out.A_indices.insert(out.A_indices.end(), row_cols[i].begin(), row_cols[i].end());
      out.A_values.insert(out.A_values.end(), row_vals[i].begin(), row_vals[i].end());
      out.A_offsets.push_back((int)out.A_indices.size());
    }
    out.constraint_bounds.assign(rhs.begin(), rhs.end());
    out.constraint_lower_bounds.resize(m);
    out.constraint_upper_bounds.resize(m);
    for (size_t i = 0; i < m; ++i) {
      const bool ranged = !ranges.empty() && ranges[i] != kInf;
      const double r    = ranged ? ranges[i] : 0.0;
      if (ranged && std::isnan(r)) parse_fail("Range value shouldn't be nan");
      double lo, hi;
      switch (row_kind[i]) {
        case 'E':
          lo = hi = rhs[i];
          if (ranged) (r < 0.0 ? lo : hi) += r;
          break;
        case 'G':
          lo = rhs[i];
          hi = ranged ? lo + std::fabs(r) : kInf;
          break;
        case 'L':
          hi = rhs[i];
          lo = ranged ? hi - std::fabs(r) : -kInf;
          break;
        default: parse_fail("Unsupported row type was passed to the Optimization Problem");
      }
      if (std::isnan(lo) || std::isnan(hi)) parse_fail("Constraint bound cannot be nan");
      out.constraint_lower_bounds[i] = lo;
      out.constraint_upper_bounds[i] = hi;
    }
    // Like the reference data model (fill_problem never calls set_row_types), the sense characters are
    // not kept: an MPS problem is carried in ranged form only.
  }
};

}  // namespace

lp_problem_t read_mps(const std::string& path, bool fixed_format)
{
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw lp_error(error_type_t::MpsFileError, "Error opening MPS file! Given path: " + path);
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, fp)) > 0)
    text.append(buf, got);
  std::fclose(fp);
  reader_t r(fixed_format);
  r.run(text);
  return std::move(r.out);
}

}  // namespace cuopt_b200
