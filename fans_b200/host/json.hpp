// json.hpp — small self-contained JSON reader for FANS input files (the reference uses nlohmann::json,
// src/reader.cpp:63-183; it is not available in this image).  Objects keep insertion order.
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fans {

class Json {
  public:
    enum Type { Null, Bool, Number, String, Array, Object };
    Type type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    bool is_null() const { return type == Null; }
    bool is_array() const { return type == Array; }
    bool is_object() const { return type == Object; }
    bool is_number() const { return type == Number; }
    bool is_string() const { return type == String; }
    bool empty() const { return type == Array ? arr.empty() : (type == Object ? obj.empty() : true); }
    size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }

    bool contains(const std::string &k) const
    {
        if (type != Object) return false;
        for (const auto &kv : obj)
            if (kv.first == k) return true;
        return false;
    }
    const Json &at(const std::string &k) const
    {
        if (type == Object)
            for (const auto &kv : obj)
                if (kv.first == k) return kv.second;
        throw std::out_of_range("key '" + k + "' not found");
    }
    const Json &operator[](const std::string &k) const { return at(k); }
    const Json &operator[](size_t i) const
    {
        if (type != Array || i >= arr.size()) throw std::out_of_range("array index out of range");
        return arr[i];
    }

    double as_double() const
    {
        if (type != Number) throw std::runtime_error("type must be number");
        return num;
    }
    int as_int() const { return (int)std::llround(as_double()); }
    bool as_bool() const
    {
        if (type != Bool) throw std::runtime_error("type must be boolean");
        return b;
    }
    const std::string &as_string() const
    {
        if (type != String) throw std::runtime_error("type must be string");
        return str;
    }
    std::vector<double> as_vector() const
    {
        if (type != Array) throw std::runtime_error("type must be array");
        std::vector<double> v;
        for (const auto &e : arr) v.push_back(e.as_double());
        return v;
    }
    std::vector<int> as_int_vector() const
    {
        if (type != Array) throw std::runtime_error("type must be array");
        std::vector<int> v;
        for (const auto &e : arr) v.push_back(e.as_int());
        return v;
    }
    std::vector<std::vector<double>> as_matrix() const
    {
        if (type != Array) throw std::runtime_error("type must be array");
        std::vector<std::vector<double>> m;
        for (const auto &e : arr) m.push_back(e.as_vector());
        return m;
    }
    std::vector<std::string> as_string_vector() const
    {
        if (type != Array) throw std::runtime_error("type must be array");
        std::vector<std::string> v;
        for (const auto &e : arr) v.push_back(e.as_string());
        return v;
    }
    double value(const std::string &k, double dflt) const { return contains(k) ? at(k).as_double() : dflt; }
    int value(const std::string &k, int dflt) const { return contains(k) ? at(k).as_int() : dflt; }
    bool value(const std::string &k, bool dflt) const { return contains(k) ? at(k).as_bool() : dflt; }

    static Json parse(const std::string &text)
    {
        size_t p = 0;
        Json j = parse_value(text, p);
        skip_ws(text, p);
        if (p != text.size()) throw std::runtime_error("JSON: trailing characters at offset " + std::to_string(p));
        return j;
    }

  private:
    static void skip_ws(const std::string &s, size_t &p)
    {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) ++p;
    }
    static Json parse_value(const std::string &s, size_t &p)
    {
        skip_ws(s, p);
        if (p >= s.size()) throw std::runtime_error("JSON: unexpected end of input");
        Json j;
        const char c = s[p];
        if (c == '{') {
            j.type = Object;
            ++p;
            skip_ws(s, p);
            if (p < s.size() && s[p] == '}') {
                ++p;
                return j;
            }
            while (true) {
                skip_ws(s, p);
                if (p >= s.size() || s[p] != '"') throw std::runtime_error("JSON: expected string key at offset " + std::to_string(p));
                std::string k = parse_string(s, p);
                skip_ws(s, p);
                if (p >= s.size() || s[p] != ':') throw std::runtime_error("JSON: expected ':' at offset " + std::to_string(p));
                ++p;
                Json v = parse_value(s, p);
                j.obj.emplace_back(std::move(k), std::move(v));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') {
                    ++p;
                    continue;
                }
                if (p < s.size() && s[p] == '}') {
                    ++p;
                    return j;
                }
                throw std::runtime_error("JSON: expected ',' or '}' at offset " + std::to_string(p));
            }
        }
        if (c == '[') {
            j.type = Array;
            ++p;
            skip_ws(s, p);
            if (p < s.size() && s[p] == ']') {
                ++p;
                return j;
            }
            while (true) {
                j.arr.push_back(parse_value(s, p));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') {
                    ++p;
                    continue;
                }
                if (p < s.size() && s[p] == ']') {
                    ++p;
                    return j;
                }
                throw std::runtime_error("JSON: expected ',' or ']' at offset " + std::to_string(p));
            }
        }
        if (c == '"') {
            j.type = String;
            j.str = parse_string(s, p);
            return j;
        }
        if (s.compare(p, 4, "true") == 0) {
            j.type = Bool, j.b = true, p += 4;
            return j;
        }
        if (s.compare(p, 5, "false") == 0) {
            j.type = Bool, j.b = false, p += 5;
            return j;
        }
        if (s.compare(p, 4, "null") == 0) {
            p += 4;
            return j;
        }
        // number
        const char *start = s.c_str() + p;
        char *end = nullptr;
        const double v = std::strtod(start, &end);
        if (end == start) throw std::runtime_error("JSON: invalid token at offset " + std::to_string(p));
        p += (size_t)(end - start);
        j.type = Number;
        j.num = v;
        return j;
    }
    static std::string parse_string(const std::string &s, size_t &p)
    {
        std::string out;
        ++p;  // opening quote
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                ++p;
                switch (s[p]) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    unsigned code = (unsigned)std::strtoul(s.substr(p + 1, 4).c_str(), nullptr, 16);
                    p += 4;
                    if (code < 0x80) out += (char)code;
                    else if (code < 0x800) {
                        out += (char)(0xC0 | (code >> 6));
                        out += (char)(0x80 | (code & 0x3F));
                    } else {
                        out += (char)(0xE0 | (code >> 12));
                        out += (char)(0x80 | ((code >> 6) & 0x3F));
                        out += (char)(0x80 | (code & 0x3F));
                    }
                    break;
                }
                default: out += s[p];
                }
                ++p;
            } else {
                out += s[p++];
            }
        }
        if (p >= s.size()) throw std::runtime_error("JSON: unterminated string");
        ++p;  // closing quote
        return out;
    }
};

}  // namespace fans
