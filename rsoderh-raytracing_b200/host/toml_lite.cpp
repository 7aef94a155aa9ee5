#include "toml_lite.hpp"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace rt {
namespace toml {
namespace {

struct Parser {
    const std::string& s;
    size_t p = 0;
    int line = 1;
    ValuePtr root = std::make_shared<Value>();
    Value* current = nullptr;

    explicit Parser(const std::string& text) : s(text) { current = root.get(); }

    [[noreturn]] void fail(const std::string& msg) const { throw ParseError(line, msg); }
    bool eof() const { return p >= s.size(); }
    char peek() const { return eof() ? '\0' : s[p]; }
    char get() {
        char c = s[p++];
        if (c == '\n') line++;
        return c;
    }
    void skip_ws() {
        while (!eof() && (peek() == ' ' || peek() == '\t')) p++;
    }
    void skip_comment() {
        if (peek() == '#')
            while (!eof() && peek() != '\n') p++;
    }
    // Skips whitespace, comments and newlines (used inside arrays).
    void skip_ws_nl() {
        for (;;) {
            skip_ws();
            if (peek() == '#') {
                skip_comment();
            } else if (peek() == '\n') {
                get();
            } else if (peek() == '\r' && p + 1 < s.size() && s[p + 1] == '\n') {
                p++;
                get();
            } else {
                return;
            }
        }
    }
    void expect_eol() {
        skip_ws();
        skip_comment();
        if (eof()) return;
        if (peek() == '\r') p++;
        if (eof()) return;
        if (peek() != '\n') fail(std::string("expected end of line, found '") + peek() + "'");
        get();
    }

    static bool bare_key_char(char c) {
        return std::isalnum((unsigned char)c) || c == '_' || c == '-';
    }

    std::string parse_basic_string() {
        // Opening quote already consumed.
        std::string out;
        for (;;) {
            if (eof()) fail("unterminated string");
            char c = get();
            if (c == '"') return out;
            if (c == '\n') fail("newline in basic string");
            if (c == '\\') {
                if (eof()) fail("unterminated escape");
                char e = get();
                switch (e) {
                    case 'b': out += '\b'; break;
                    case 't': out += '\t'; break;
                    case 'n': out += '\n'; break;
                    case 'f': out += '\f'; break;
                    case 'r': out += '\r'; break;
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case 'u':
                    case 'U': {
                        int n = e == 'u' ? 4 : 8;
                        uint32_t cp = 0;
                        for (int i = 0; i < n; i++) {
                            if (eof()) fail("unterminated unicode escape");
                            char h = get();
                            cp <<= 4;
                            if (h >= '0' && h <= '9') cp |= h - '0';
                            else if (h >= 'a' && h <= 'f') cp |= h - 'a' + 10;
                            else if (h >= 'A' && h <= 'F') cp |= h - 'A' + 10;
                            else fail("bad unicode escape");
                        }
                        // UTF-8 encode.
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) {
                            out += (char)(0xC0 | (cp >> 6));
                            out += (char)(0x80 | (cp & 0x3F));
                        } else if (cp < 0x10000) {
                            out += (char)(0xE0 | (cp >> 12));
                            out += (char)(0x80 | ((cp >> 6) & 0x3F));
                            out += (char)(0x80 | (cp & 0x3F));
                        } else {
                            out += (char)(0xF0 | (cp >> 18));
                            out += (char)(0x80 | ((cp >> 12) & 0x3F));
                            out += (char)(0x80 | ((cp >> 6) & 0x3F));
                            out += (char)(0x80 | (cp & 0x3F));
                        }
                        break;
                    }
                    default: fail(std::string("unknown escape \\") + e);
                }
            } else {
                out += c;
            }
        }
    }
    std::string parse_literal_string() {
        std::string out;
        for (;;) {
            if (eof()) fail("unterminated literal string");
            char c = get();
            if (c == '\'') return out;
            if (c == '\n') fail("newline in literal string");
            out += c;
        }
    }

    std::string parse_key_part() {
        skip_ws();
        if (peek() == '"') {
            get();
            return parse_basic_string();
        }
        if (peek() == '\'') {
            get();
            return parse_literal_string();
        }
        std::string out;
        while (!eof() && bare_key_char(peek())) out += get();
        if (out.empty()) fail(std::string("expected a key, found '") + peek() + "'");
        return out;
    }
    std::vector<std::string> parse_dotted_key() {
        std::vector<std::string> parts;
        parts.push_back(parse_key_part());
        skip_ws();
        while (peek() == '.') {
            get();
            parts.push_back(parse_key_part());
            skip_ws();
        }
        return parts;
    }

    ValuePtr parse_number_or_bool() {
        size_t start = p;
        while (!eof()) {
            char c = peek();
            if (std::isalnum((unsigned char)c) || c == '+' || c == '-' || c == '.' || c == '_' ||
                c == ':')
                p++;
            else
                break;
        }
        std::string tok = s.substr(start, p - start);
        if (tok.empty()) fail(std::string("expected a value, found '") + peek() + "'");
        auto v = std::make_shared<Value>();
        v->line = line;
        if (tok == "true" || tok == "false") {
            v->kind = Value::Boolean;
            v->boolean = tok == "true";
            return v;
        }
        std::string clean;
        for (char c : tok)
            if (c != '_') clean += c;
        std::string unsign = clean;
        double sign = 1.0;
        if (!unsign.empty() && (unsign[0] == '+' || unsign[0] == '-')) {
            if (unsign[0] == '-') sign = -1.0;
            unsign = unsign.substr(1);
        }
        if (unsign == "inf") {
            v->kind = Value::Float;
            v->real = sign * INFINITY;
            return v;
        }
        if (unsign == "nan") {
            v->kind = Value::Float;
            v->real = NAN;
            return v;
        }
        if (unsign.empty() || !std::isdigit((unsigned char)unsign[0]))
            fail("invalid value '" + tok + "'");
        bool is_float = false;
        bool is_hex = unsign.size() > 2 && unsign[0] == '0' &&
                      (unsign[1] == 'x' || unsign[1] == 'o' || unsign[1] == 'b');
        if (!is_hex)
            for (char c : unsign)
                if (c == '.' || c == 'e' || c == 'E') is_float = true;
        char* end = nullptr;
        if (is_float) {
            v->kind = Value::Float;
            v->real = std::strtod(clean.c_str(), &end);
            if (!end || *end != '\0') fail("invalid float '" + tok + "'");
        } else {
            v->kind = Value::Integer;
            if (is_hex) {
                int base = unsign[1] == 'x' ? 16 : unsign[1] == 'o' ? 8 : 2;
                v->integer =
                    (int64_t)(sign * (double)std::strtoull(unsign.c_str() + 2, &end, base));
            } else {
                v->integer = std::strtoll(clean.c_str(), &end, 10);
            }
            if (!end || *end != '\0') fail("invalid integer '" + tok + "'");
        }
        return v;
    }

    ValuePtr parse_value() {
        skip_ws();
        char c = peek();
        auto v = std::make_shared<Value>();
        v->line = line;
        if (c == '"') {
            if (s.compare(p, 3, "\"\"\"") == 0) fail("multi-line strings are not supported");
            get();
            v->kind = Value::String;
            v->str = parse_basic_string();
            return v;
        }
        if (c == '\'') {
            if (s.compare(p, 3, "'''") == 0) fail("multi-line strings are not supported");
            get();
            v->kind = Value::String;
            v->str = parse_literal_string();
            return v;
        }
        if (c == '[') {
            get();
            v->kind = Value::Array;
            for (;;) {
                skip_ws_nl();
                if (peek() == ']') {
                    get();
                    break;
                }
                v->array.push_back(parse_value());
                skip_ws_nl();
                if (peek() == ',') {
                    get();
                    continue;
                }
                if (peek() == ']') {
                    get();
                    break;
                }
                fail("expected ',' or ']' in array");
            }
            return v;
        }
        if (c == '{') {
            get();
            v->kind = Value::Table;
            v->defined_inline = true;
            skip_ws();
            if (peek() == '}') {
                get();
                return v;
            }
            for (;;) {
                auto key = parse_dotted_key();
                skip_ws();
                if (peek() != '=') fail("expected '=' in inline table");
                get();
                auto val = parse_value();
                insert(v.get(), key, val);
                skip_ws();
                if (peek() == ',') {
                    get();
                    continue;
                }
                if (peek() == '}') {
                    get();
                    break;
                }
                fail("expected ',' or '}' in inline table");
            }
            return v;
        }
        return parse_number_or_bool();
    }

    // Descends (creating tables) along all but the last key part, then inserts.
    void insert(Value* table, const std::vector<std::string>& key, ValuePtr val) {
        Value* t = table;
        for (size_t i = 0; i + 1 < key.size(); i++) {
            Value* next = t->find_mut(key[i]);
            if (!next) {
                auto nt = std::make_shared<Value>();
                nt->kind = Value::Table;
                nt->line = line;
                t->table.emplace_back(key[i], nt);
                next = nt.get();
            } else if (next->kind != Value::Table) {
                fail("key '" + key[i] + "' is not a table");
            }
            t = next;
        }
        if (t->find(key.back())) fail("duplicate key '" + key.back() + "'");
        t->table.emplace_back(key.back(), val);
    }

    // Resolves a header path. For [[a.b]] the last part appends a new table to an array.
    Value* resolve_header(const std::vector<std::string>& path, bool array_header) {
        Value* t = root.get();
        for (size_t i = 0; i < path.size(); i++) {
            bool last = i + 1 == path.size();
            Value* next = t->find_mut(path[i]);
            if (last && array_header) {
                if (!next) {
                    auto arr = std::make_shared<Value>();
                    arr->kind = Value::Array;
                    arr->is_array_of_tables = true;
                    arr->line = line;
                    t->table.emplace_back(path[i], arr);
                    next = arr.get();
                } else if (!(next->kind == Value::Array && next->is_array_of_tables)) {
                    fail("'" + path[i] + "' is not an array of tables");
                }
                auto elem = std::make_shared<Value>();
                elem->kind = Value::Table;
                elem->line = line;
                next->array.push_back(elem);
                return elem.get();
            }
            if (!next) {
                auto nt = std::make_shared<Value>();
                nt->kind = Value::Table;
                nt->line = line;
                t->table.emplace_back(path[i], nt);
                next = nt.get();
            }
            if (next->kind == Value::Array && next->is_array_of_tables) {
                if (next->array.empty()) fail("empty array of tables '" + path[i] + "'");
                next = next->array.back().get();
            } else if (next->kind != Value::Table || next->defined_inline) {
                fail("'" + path[i] + "' is not a table");
            }
            t = next;
        }
        return t;
    }

    ValuePtr run() {
        while (!eof()) {
            skip_ws();
            char c = peek();
            if (eof()) break;
            if (c == '#') {
                skip_comment();
                continue;
            }
            if (c == '\n') {
                get();
                continue;
            }
            if (c == '\r') {
                p++;
                continue;
            }
            if (c == '[') {
                get();
                bool array_header = false;
                if (peek() == '[') {
                    get();
                    array_header = true;
                }
                auto path = parse_dotted_key();
                skip_ws();
                if (peek() != ']') fail("expected ']' to close table header");
                get();
                if (array_header) {
                    if (peek() != ']') fail("expected ']]' to close array-of-tables header");
                    get();
                }
                current = resolve_header(path, array_header);
                expect_eol();
                continue;
            }
            auto key = parse_dotted_key();
            skip_ws();
            if (peek() != '=') fail("expected '=' after key '" + key.back() + "'");
            get();
            auto val = parse_value();
            insert(current, key, val);
            expect_eol();
        }
        return root;
    }
};

}  // namespace

ValuePtr parse(const std::string& text) {
    Parser parser(text);
    return parser.run();
}

}  // namespace toml
}  // namespace rt
