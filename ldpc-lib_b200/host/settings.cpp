#include "settings.h"

#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>

namespace {

bool read_whole_file(std::string const& name, std::string& out)
{
    std::ifstream f(name.c_str(), std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

std::string directory_of(std::string const& filename)
{
    size_t k = filename.rfind('/');
    if (k == std::string::npos) k = filename.rfind('\\');
    return k == std::string::npos ? std::string(".") : filename.substr(0, k);
}

bool identifier_char(char v) { return v == '_' || (v >= 'a' && v <= 'z') || (v >= 'A' && v <= 'Z') || (v >= '0' && v <= '9') || v == '-'; }
bool number_char(char v) { return v == '.' || (v >= '0' && v <= '9') || v == '-' || v == 'e' || v == 'E'; }

} // namespace

// a position in an in-memory copy of one file
struct settings::cursor {
    std::string const& s;
    size_t i;
    explicit cursor(std::string const& text) : s(text), i(0) {}
    bool eof() const { return i >= s.size(); }
    char peek() const { return s[i]; }
    void skip_ws()
    {
        for (;;) {
            while (!eof() && (unsigned char)peek() <= ' ') ++i;
            if (!eof() && peek() == '/') { while (!eof() && s[i++] != '\n') {} }
            else break;
        }
    }
    void expect(std::string const& text, std::string const& path, std::string const& prefix)
    {
        std::string seen;
        for (size_t k = 0; k < text.size(); ++k) {
            if (eof()) die("Unexpected end of file where '%s' expected (current context: '%s%s')", text.c_str(), prefix.c_str(), path.c_str());
            char ch = s[i++];
            seen.push_back(ch);
            if (ch != text[k])
                die("Expected '%s' found '%s%s' (current context: '%s%s')", text.c_str(), seen.c_str(), text.size() == 1 ? "" : "...", prefix.c_str(), path.c_str());
        }
    }
    std::string quoted(std::string const& path, std::string const& prefix)
    {
        expect("\"", path, prefix);
        std::string buf;
        while (!eof() && peek() != '"') buf += s[i++];
        expect("\"", path, prefix);
        return buf;
    }
    std::string run(bool (*pred)(char))
    {
        std::string buf;
        while (!eof() && pred(peek())) buf += s[i++];
        return buf;
    }
};

settings::settings() : type_(RECORD), mrows_(0), mcols_(0) {}

void settings::clear()
{
    str_.clear(); vec_.clear(); rec_.clear(); mrows_ = mcols_ = 0;
}

settings settings::from_file(std::string const& filename)
{
    std::string text;
    if (!read_whole_file(filename, text)) die("File not found: %s", filename.c_str());
    cursor in(text);
    settings s;
    s.parse(in, directory_of(filename), "", filename);
    return s;
}

settings settings::from_string(std::string const& text, std::string const& current_directory)
{
    cursor in(text);
    settings s;
    s.parse(in, current_directory, "", "");
    return s;
}

void settings::parse(cursor& in, std::string const& dir, std::string const& path, std::string const& prefix)
{
    clear();
    path_ = path; path_prefix_ = prefix;
    in.skip_ws();
    if (in.eof()) { type_ = RECORD; return; }            // an empty stream is an empty record
    auto dims = [&](int& n_rows, int& n_cols) {
        in.skip_ws();
        in.expect("(", path, prefix);
        settings r, c;
        r.parse(in, dir, path + "/#rows", prefix); r.cast_to(n_rows);
        c.parse(in, dir, path + "/#cols", prefix); c.cast_to(n_cols);
        in.skip_ws();
        in.expect(")", path, prefix);
        if (n_rows < 0 || n_cols < 0) die("Negative matrix size (current context: '%s%s')", prefix.c_str(), path.c_str());
        in.skip_ws();
        in.expect("{", path, prefix);
    };
    auto child_path = [&](size_t k) { std::ostringstream o; o << path << "/" << k; return o.str(); };
    switch (in.peek()) {
    case '"':
        type_ = STRING;
        str_ = in.quoted(path, prefix);
        break;
    case 'a': {
        type_ = ARRAY;
        in.expect("array", path, prefix);
        in.skip_ws();
        if (!in.eof() && in.peek() == '{') {
            in.expect("{", path, prefix);
            in.skip_ws();
            while (!in.eof() && in.peek() != '}') {
                size_t before = in.i;
                settings el;
                el.parse(in, dir, child_path(vec_.size()), prefix);
                if (in.i == before) die("Unexpected character '%c' in array (current context: '%s%s')", in.peek(), prefix.c_str(), path.c_str());
                vec_.push_back(el);
                in.skip_ws();
            }
            in.expect("}", path, prefix);
        } else {
            in.expect("@", path, prefix);
            in.skip_ws();
            std::string file = dir + "/" + in.quoted(path, prefix), text;
            read_whole_file(file, text);                   // a missing file is an empty array, as in the reference
            cursor sub(text);
            std::string subdir = directory_of(file);
            sub.skip_ws();
            while (!sub.eof()) {
                size_t before = sub.i;
                settings el;
                el.parse(sub, subdir, child_path(vec_.size()), prefix);
                if (sub.i == before) die("Unexpected character '%c' in %s", sub.peek(), file.c_str());
                vec_.push_back(el);
                sub.skip_ws();
            }
        }
        break;
    }
    case 'm': {
        type_ = MATRIX;
        in.expect("matrix", path, prefix);
        dims(mrows_, mcols_);
        vec_.resize((size_t)mrows_ * mcols_);
        for (int r = 0; r < mrows_; ++r)
            for (int c = 0; c < mcols_; ++c) {
                std::ostringstream o; o << "/(" << r << "," << c << ")";
                vec_[(size_t)r * mcols_ + c].parse(in, dir, path + o.str(), prefix);
            }
        in.skip_ws();
        in.expect("}", path, prefix);
        break;
    }
    case 's': {
        type_ = MATRIX;
        in.expect("sparse", path, prefix);
        in.skip_ws();
        in.expect("matrix", path, prefix);
        dims(mrows_, mcols_);
        vec_.assign((size_t)mrows_ * mcols_, settings());
        in.skip_ws();
        int count = 0;
        while (!in.eof() && in.peek() != '}') {
            std::ostringstream xr, xc, xv;
            xr << path << "/row#" << count; xc << path << "/col#" << count; xv << path << "/val#" << count;
            int row = -1, col = -1;
            size_t before = in.i;
            settings t;
            t.parse(in, dir, xr.str(), prefix); t.cast_to(row);
            if (in.i == before) die("Unexpected character '%c' in sparse matrix (current context: '%s%s')", in.peek(), prefix.c_str(), path.c_str());
            if (row < 0 || row >= mrows_) die("Row index out of bounds: %d, number of rows: %d (current context: '%s%s')", row, mrows_, prefix.c_str(), xr.str().c_str());
            t.parse(in, dir, xc.str(), prefix); t.cast_to(col);
            if (col < 0 || col >= mcols_) die("Column index out of bounds: %d, number of columns: %d (current context: '%s%s')", col, mcols_, prefix.c_str(), xc.str().c_str());
            vec_[(size_t)row * mcols_ + col].parse(in, dir, xv.str(), prefix);
            in.skip_ws();
            ++count;
        }
        in.expect("}", path, prefix);
        break;
    }
    case '{': {
        type_ = RECORD;
        in.expect("{", path, prefix);
        in.skip_ws();
        while (!in.eof() && in.peek() != '}') {
            std::string id = in.run(identifier_char);
            in.skip_ws();
            in.expect("=", path, prefix);
            settings field;
            field.parse(in, dir, path + "/" + id, prefix);
            in.skip_ws();
            rec_.insert(std::make_pair(id, field));        // the first definition of a key wins
        }
        in.expect("}", path, prefix);
        break;
    }
    case '@': {
        in.expect("@", path, prefix);
        std::string file = dir + "/" + in.quoted(path, prefix), text;
        if (!read_whole_file(file, text)) die("Stream is invalid. Current directory = '%s', path = '%s', path_prefix = '%s'", directory_of(file).c_str(), path.c_str(), prefix.c_str());
        cursor sub(text);
        parse(sub, directory_of(file), path, prefix);
        break;
    }
    default:
        type_ = STRING;
        str_ = in.run(number_char);
        break;
    }
}

bool settings::can_select(std::string const& path) const
{
    if (path.empty()) return true;
    if (type_ != RECORD) return false;
    size_t slash = path.find('/');
    std::string key = path.substr(0, slash);
    std::string rest = slash == std::string::npos ? "" : path.substr(slash + 1);
    auto it = rec_.find(key);
    if (it != rec_.end()) return it->second.can_select(rest);
    auto def = rec_.find("defaults");
    return def != rec_.end() && def->second.can_select(path);
}

settings const& settings::select(std::string const& path) const
{
    if (path.empty()) return *this;
    if (type_ != RECORD) die("Object '%s' is not a record, it is impossible to select from it!", get_full_path().c_str());
    size_t slash = path.find('/');
    std::string key = path.substr(0, slash);
    std::string rest = slash == std::string::npos ? "" : path.substr(slash + 1);
    auto it = rec_.find(key);
    if (it != rec_.end()) return it->second.select(rest);
    auto def = rec_.find("defaults");
    if (def != rec_.end()) {
        if (def->second.can_select(path)) return def->second.select(path);
        die("Object '%s' does not contain a field '%s' ('defaults' has also been checked)!", get_full_path().c_str(), key.c_str());
    }
    die("Object '%s' does not contain a field '%s'!", get_full_path().c_str(), key.c_str());
}

settings& settings::open(std::string const& path)
{
    if (path.empty()) return *this;
    if (type_ != RECORD) die("Object '%s' is not a record, it is impossible to open it!", get_full_path().c_str());
    size_t slash = path.find('/');
    std::string key = path.substr(0, slash);
    std::string rest = slash == std::string::npos ? "" : path.substr(slash + 1);
    auto it = rec_.find(key);
    if (it != rec_.end()) return it->second.open(rest);
    settings& created = rec_[key];
    created.path_prefix_ = path_prefix_;
    created.path_ = path_ + "/" + key;
    return created.open(rest);
}

void settings::print(std::ostream& stream, int indent, bool newline) const
{
    const char last = newline ? '\n' : ' ';
    auto pad = [](int n) { return std::string((size_t)n * 2, ' '); };
    switch (type_) {
    case RECORD:
        stream << "{\n";
        for (auto const& kv : rec_) {
            stream << pad(indent + 1) << kv.first << " = ";
            kv.second.print(stream, indent + 1, true);
        }
        stream << pad(indent) << "}" << last;
        break;
    case STRING: {
        bool numeric = true;
        for (char ch : str_) numeric &= number_char(ch);
        if (numeric) stream << str_ << last;
        else stream << '"' << str_ << '"' << last;
        break;
    }
    case ARRAY: {
        bool nested = false;
        for (auto const& el : vec_) nested |= el.type_ != STRING;
        stream << "array {" << (nested ? "\n" : " ");
        for (auto const& el : vec_) {
            if (nested) stream << pad(indent + 1);
            el.print(stream, indent + 1, nested);
        }
        if (nested) stream << pad(indent);
        stream << "}" << last;
        break;
    }
    case MATRIX: {
        stream << "matrix (" << mrows_ << " " << mcols_ << ") {\n";
        size_t width = 0;
        for (auto const& cell : vec_) {
            std::ostringstream oss;
            cell.print(oss, indent + 1, false);
            size_t len = oss.str().size();
            if (len > 0 && len - 1 > width) width = len - 1;
        }
        for (int r = 0; r < mrows_; ++r) {
            stream << pad(indent + 1);
            for (int c = 0; c < mcols_; ++c) {
                stream << std::setw((int)width);
                vec_[(size_t)r * mcols_ + c].print(stream, indent + 1, c + 1 == mcols_);
            }
        }
        stream << pad(indent) << "}" << last;
        break;
    }
    case OBJECT_REF:
        stream << "@\"" << str_ << '"' << last;
        break;
    case ARRAY_REF:
        stream << "array @\"" << str_ << '"' << last;
        break;
    }
}

void settings::to_stream(std::ostream& stream) const { print(stream, 0, true); }

std::string settings::to_string() const
{
    std::ostringstream oss;
    to_stream(oss);
    return oss.str();
}

void settings::to_file(std::string const& filename, bool append) const
{
    std::ofstream out(filename.c_str(), append ? std::ios_base::app : std::ios_base::out);
    to_stream(out);
}

void settings::set(settings const& value)
{
    switch (value.type_) {
    case RECORD: {
        clear();
        type_ = RECORD;
        for (auto const& kv : value.rec_) open(kv.first).set(kv.second);
        break;
    }
    case STRING: set(value.str_); break;
    case ARRAY: set(value.vec_); break;
    case MATRIX: {
        clear();
        type_ = MATRIX;
        mrows_ = value.mrows_; mcols_ = value.mcols_;
        vec_.resize(value.vec_.size());
        for (size_t i = 0; i < vec_.size(); ++i) vec_[i].set(value.vec_[i]);
        break;
    }
    case OBJECT_REF: set_object_ref(value.str_); break;
    case ARRAY_REF: set_array_ref(value.str_); break;
    }
}

void settings::cast_to(std::string& target) const
{
    if (type_ != STRING) die("Cannot convert a non-string object '%s' into a string!", get_full_path().c_str());
    target = str_;
}

void settings::cast_to(char* target) const
{
    if (type_ != STRING) die("Cannot convert a non-string object '%s' into what you ask!", get_full_path().c_str());
    std::istringstream iss(str_);
    std::string word;
    iss >> word;                                            // operator>>(char*) semantics: one whitespace-free token
    snprintf(target, 512, "%s", word.c_str());
}

void settings::set_object_ref(std::string const& path) { clear(); type_ = OBJECT_REF; str_ = path; }
void settings::set_array_ref(std::string const& path) { clear(); type_ = ARRAY_REF; str_ = path; }
