// The .jsonx configuration dialect of ldpc-lib: parser, accessor and pretty-printer.
//
// Drop-in for the reference's `settings` class (settings.h:26-160, settings.cpp) on everything the
// simulation path uses: same class and method names, same grammar, same lookup rules and the same
// output text, so inputs keep parsing identically and result files stay interchangeable.
//
//   object = number | "string" | array { object* } | array @"file" | matrix (r c) { object* }
//          | sparse matrix (r c) { [row col object]* } | { [identifier = object]* } | @"file"
//
// Behaviours kept on purpose (each is observable through the files in /files and /files/tests):
//   * '/' starts a comment that runs to the end of the line (settings.cpp:111-122)
//   * a key that appears twice in a record keeps its FIRST value (std::map::insert, settings.cpp:315)
//   * select()/can_select() fall back to the record's `defaults` member, recursively, with the WHOLE
//     remaining path (settings.cpp:341-358, 364-399)
//   * `@"file"` and `array @"file"` resolve relative to the directory of the file that names them
//   * records print their keys in sorted order, arrays of plain values print on one line, matrices
//     right-align their cells to the widest one (settings.cpp:431-505)
#pragma once
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "commons.h"

class settings {
public:
    enum object_type { RECORD, STRING, ARRAY, MATRIX, OBJECT_REF, ARRAY_REF };

    settings();                                              // an empty record "{}"
    static settings from_file(std::string const& filename);
    static settings from_string(std::string const& text, std::string const& current_directory = ".");

    void to_stream(std::ostream& stream) const;
    void to_file(std::string const& filename, bool append = false) const;
    std::string to_string() const;

    std::string get_full_path() const { return path_prefix_ + path_; }
    std::string get_path() const { return path_; }
    object_type get_type() const { return type_; }

    bool can_select(std::string const& path) const;
    settings const& select(std::string const& path) const;
    settings& open(std::string const& path);                 // creates missing records on the way

    template <typename result_t> void cast_to(result_t& target) const
    {
        if (type_ != STRING) die("Cannot convert a non-string object '%s' into what you ask!", get_full_path().c_str());
        std::istringstream iss(str_);
        iss >> target;
    }
    void cast_to(std::string& target) const;
    void cast_to(char* target) const;                        // `char mark_file[512]` in main_simulation.cpp:239
    void cast_to(settings& target) const { target = *this; }
    template <typename element_t> void cast_to(std::vector<element_t>& target) const
    {
        if (type_ != ARRAY) die("Cannot convert a non-array object '%s' into a vector!", get_full_path().c_str());
        std::vector<element_t> rv(vec_.size());
        for (size_t i = 0; i < vec_.size(); ++i) vec_[i].cast_to(rv[i]);
        target = rv;
    }
    template <typename element_t> void cast_to(matrix<element_t>& target) const
    {
        if (type_ != MATRIX) die("Cannot convert a non-matrix object '%s' into a matrix!", get_full_path().c_str());
        matrix<element_t> rv(mrows_, mcols_);
        for (int r = 0; r < mrows_; ++r)
            for (int c = 0; c < mcols_; ++c) {
                settings const& cell = vec_[(size_t)r * mcols_ + c];
                if (cell.get_path().empty()) rv(r, c) = element_t();     // unset cell of a sparse matrix
                else cell.cast_to(rv(r, c));
            }
        target = rv;
    }

    void set(settings const& value);
    void set_object_ref(std::string const& path);
    void set_array_ref(std::string const& path);
    template <typename value_t> void set(value_t const& value)
    {
        clear();
        type_ = STRING;
        std::ostringstream oss;
        oss << value;
        str_ = oss.str();
    }
    template <typename value_t> void set(std::vector<value_t> const& value)
    {
        clear();
        type_ = ARRAY;
        vec_.resize(value.size());
        for (size_t i = 0; i < value.size(); ++i) vec_[i].set(value[i]);
    }
    template <typename value_t> void set(matrix<value_t> const& value)
    {
        clear();
        type_ = MATRIX;
        mrows_ = value.n_rows(); mcols_ = value.n_cols();
        vec_.resize((size_t)mrows_ * mcols_);
        for (int r = 0; r < mrows_; ++r)
            for (int c = 0; c < mcols_; ++c) vec_[(size_t)r * mcols_ + c].set(value(r, c));
    }
    template <typename value_t> void append(value_t const& value)
    {
        if (type_ != ARRAY) die("Cannot append to a non-array object '%s'!", get_full_path().c_str());
        settings nv;
        nv.set(value);
        vec_.push_back(nv);
    }

private:
    object_type type_;
    std::string path_, path_prefix_;
    std::string str_;                        // STRING, OBJECT_REF, ARRAY_REF
    std::vector<settings> vec_;              // ARRAY elements, MATRIX cells (row major)
    int mrows_, mcols_;
    std::map<std::string, settings> rec_;    // RECORD

    struct cursor;
    void clear();
    void parse(cursor& in, std::string const& dir, std::string const& path, std::string const& prefix);
    void print(std::ostream& stream, int indent, bool newline) const;
};
