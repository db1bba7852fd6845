// tests/cxx/boost_mini -- TEST INFRASTRUCTURE ONLY (see ptree.hpp): read_json / write_json for the JSON dialect of
// DROPS' parameter files (objects, arrays, strings, numbers, true/false/null, // and /* */ comments).  As in
// Boost.PropertyTree every scalar is kept as its text, arrays become children with empty keys.
#ifndef DROPS_B200_TESTS_BOOST_MINI_JSON_PARSER_HPP
#define DROPS_B200_TESTS_BOOST_MINI_JSON_PARSER_HPP

#include <boost/property_tree/ptree.hpp>

#include <cctype>
#include <fstream>
#include <istream>
#include <iterator>
#include <ostream>

namespace boost { namespace property_tree {

class json_parser_error : public ptree_error {
  public:
    explicit json_parser_error(const std::string& w) : ptree_error(w) {}
};

namespace json_mini {

struct Reader {
    const std::string& s;
    size_t i;
    explicit Reader(const std::string& text) : s(text), i(0) {}

    [[noreturn]] void fail(const std::string& what) const
    {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); ++k) if (s[k] == '\n') ++line;
        std::ostringstream os;
        os << "json: " << what << " (line " << line << ")";
        throw json_parser_error(os.str());
    }
    void skip()
    {
        for (;;) {
            while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
            if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') { while (i < s.size() && s[i] != '\n') ++i; continue; }
            if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
                const size_t e = s.find("*/", i + 2);
                if (e == std::string::npos) fail("unterminated comment");
                i = e + 2;
                continue;
            }
            return;
        }
    }
    std::string string_()
    {
        if (s[i] != '"') fail("expected string");
        ++i;
        std::string out;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char c = s[i + 1];
                i += 2;
                switch (c) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (i + 4 > s.size()) fail("bad \\u escape");
                        const unsigned cp = static_cast<unsigned>(std::stoul(s.substr(i, 4), nullptr, 16));
                        i += 4;
                        if (cp < 0x80) out += static_cast<char>(cp);
                        else if (cp < 0x800) { out += static_cast<char>(0xC0 | (cp >> 6)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                        else { out += static_cast<char>(0xE0 | (cp >> 12)); out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += c;
                }
            } else out += s[i++];
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    void value(ptree& node)
    {
        skip();
        if (i >= s.size()) fail("unexpected end of input");
        const char c = s[i];
        if (c == '{') {
            ++i; skip();
            if (i < s.size() && s[i] == '}') { ++i; return; }
            for (;;) {
                skip();
                const std::string key = string_();
                skip();
                if (i >= s.size() || s[i] != ':') fail("expected ':'");
                ++i;
                ptree child;
                value(child);
                node.push_back(ptree::value_type(key, child));
                skip();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; return; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            ++i; skip();
            if (i < s.size() && s[i] == ']') { ++i; return; }
            for (;;) {
                ptree child;
                value(child);
                node.push_back(ptree::value_type(std::string(), child));
                skip();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; return; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            node.data() = string_();
        } else {
            const size_t b = i;
            while (i < s.size() && (std::isalnum(static_cast<unsigned char>(s[i])) || s[i] == '+' || s[i] == '-' || s[i] == '.')) ++i;
            if (i == b) fail("unexpected character");
            node.data() = s.substr(b, i - b);
        }
    }
};

inline void write_string(std::ostream& os, const std::string& v)
{
    os << '"';
    for (char c : v) {
        if (c == '"' || c == '\\') os << '\\' << c;
        else if (c == '\n') os << "\\n";
        else if (c == '\t') os << "\\t";
        else os << c;
    }
    os << '"';
}
inline void write(std::ostream& os, const ptree& t, int indent)
{
    const std::string pad(static_cast<size_t>(4 * (indent + 1)), ' ');
    if (t.empty()) { write_string(os, t.data()); return; }
    bool array = true;
    for (const ptree::value_type& c : t) if (!c.first.empty()) array = false;
    os << (array ? "[\n" : "{\n");
    size_t k = 0;
    for (const ptree::value_type& c : t) {
        os << pad;
        if (!array) { write_string(os, c.first); os << ": "; }
        write(os, c.second, indent + 1);
        os << (++k < t.size() ? ",\n" : "\n");
    }
    os << std::string(static_cast<size_t>(4 * indent), ' ') << (array ? "]" : "}");
}

}  // namespace json_mini

inline void read_json(std::istream& in, ptree& pt)
{
    const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    json_mini::Reader r(text);
    ptree out;
    r.value(out);
    r.skip();
    if (r.i != text.size()) r.fail("garbage after the top-level value");
    pt = out;
}
inline void read_json(const std::string& filename, ptree& pt)
{
    std::ifstream in(filename.c_str());
    if (!in) throw json_parser_error("cannot open file " + filename);
    read_json(in, pt);
}
inline void write_json(std::ostream& os, const ptree& pt, bool = true)
{
    json_mini::write(os, pt, 0);
    os << '\n';
}
inline void write_json(const std::string& filename, const ptree& pt)
{
    std::ofstream os(filename.c_str());
    if (!os) throw json_parser_error("cannot open file " + filename);
    write_json(os, pt);
}

}}  // namespace boost::property_tree

#endif
