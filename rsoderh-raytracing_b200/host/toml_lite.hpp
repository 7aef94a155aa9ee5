// toml_lite.hpp — a small TOML reader covering what scene files use (and a bit more):
// comments, bare/quoted/dotted keys, [table] and [[array-of-table]] headers with
// dotted paths, basic/literal strings, integers, floats, booleans, (multi-line)
// arrays and inline tables. Stands in for the `toml` 0.9.11 crate at
// src/scene.rs:243 (numbers are parsed as i64/f64 and cast by the consumer).
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rt {
namespace toml {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
    enum Kind { String, Integer, Float, Boolean, Array, Table } kind = Table;
    std::string str;
    int64_t integer = 0;
    double real = 0.0;
    bool boolean = false;
    std::vector<ValuePtr> array;
    // Insertion-ordered table.
    std::vector<std::pair<std::string, ValuePtr>> table;
    // Bookkeeping for header semantics.
    bool is_array_of_tables = false;
    bool defined_inline = false;
    int line = 0;

    const Value* find(const std::string& key) const {
        for (auto& kv : table)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    Value* find_mut(const std::string& key) {
        for (auto& kv : table)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    bool is_number() const { return kind == Integer || kind == Float; }
    double as_double() const { return kind == Integer ? (double)integer : real; }
    const char* kind_name() const {
        switch (kind) {
            case String: return "string";
            case Integer: return "integer";
            case Float: return "float";
            case Boolean: return "boolean";
            case Array: return "array";
            case Table: return "table";
        }
        return "?";
    }
};

struct ParseError : std::runtime_error {
    int line;
    ParseError(int line_, const std::string& msg)
        : std::runtime_error("TOML parse error at line " + std::to_string(line_) + ": " + msg),
          line(line_) {}
};

// Parses a whole document; throws ParseError.
ValuePtr parse(const std::string& text);

}  // namespace toml
}  // namespace rt
