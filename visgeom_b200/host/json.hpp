// json.hpp -- minimal JSON reader for the calibration problem files (the reference reads them with
// Boost.PropertyTree, include/json.h:27-84; neither Boost nor any JSON library is available here).
// Numbers are doubles, objects keep their key order, get("a.b") walks nested objects like ptree::get.
#pragma once

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace visgeom_b200 {
namespace json {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value *find(const std::string &path) const
    {
        const Value *v = this;
        size_t pos = 0;
        while (pos <= path.size()) {
            const size_t dot = path.find('.', pos);
            const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            if (v->type != Object) return nullptr;
            const Value *next = nullptr;
            for (const auto &kv : v->obj) if (kv.first == key) { next = &kv.second; break; }
            if (!next) return nullptr;
            v = next;
            if (dot == std::string::npos) break;
            pos = dot + 1;
        }
        return v;
    }
    const Value &child(const std::string &path) const
    {
        const Value *v = find(path);
        if (!v) throw std::runtime_error("No such node (" + path + ")");   // ptree_bad_path's message
        return *v;
    }
    double getDouble(const std::string &path) const
    {
        const Value &v = child(path);
        if (v.type == Number) return v.num;
        if (v.type == String) return std::strtod(v.str.c_str(), nullptr);
        throw std::runtime_error("conversion of data to type \"double\" failed (" + path + ")");
    }
    int getInt(const std::string &path) const { return (int)getDouble(path); }
    bool getBool(const std::string &path) const
    {
        const Value &v = child(path);
        if (v.type == Bool) return v.b;
        if (v.type == Number) return v.num != 0;
        if (v.type == String) return v.str == "true" || v.str == "1";
        throw std::runtime_error("conversion of data to type \"bool\" failed (" + path + ")");
    }
    std::string getString(const std::string &path) const
    {
        const Value &v = child(path);
        if (v.type != String) throw std::runtime_error("conversion of data to type \"string\" failed (" + path + ")");
        return v.str;
    }
    // numbers of an array node (readVector<double>, json.h:72-80)
    std::vector<double> numbers() const
    {
        std::vector<double> out;
        for (const Value &x : arr) out.push_back(x.type == Number ? x.num : std::strtod(x.str.c_str(), nullptr));
        return out;
    }
};

class Parser {
public:
    explicit Parser(const std::string &text) : s(text) {}
    Value parse()
    {
        Value v = value();
        ws();
        if (i != s.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string &s;
    size_t i = 0;
    [[noreturn]] void fail(const char *what) const
    {
        throw std::runtime_error(std::string("JSON: ") + what + " at offset " + std::to_string(i));
    }
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
    Value value()
    {
        ws();
        if (i >= s.size()) fail("unexpected end");
        const char c = s[i];
        Value v;
        if (c == '{') {
            v.type = Value::Object; i++; ws();
            if (i < s.size() && s[i] == '}') { i++; return v; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("expected a key");
                std::string key = string();
                ws();
                if (i >= s.size() || s[i] != ':') fail("expected ':'");
                i++;
                v.obj.emplace_back(std::move(key), value());
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == '}') { i++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = Value::Array; i++; ws();
            if (i < s.size() && s[i] == ']') { i++; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == ']') { i++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = Value::String; v.str = string();
        } else if (s.compare(i, 4, "true") == 0) { v.type = Value::Bool; v.b = true; i += 4; }
        else if (s.compare(i, 5, "false") == 0) { v.type = Value::Bool; v.b = false; i += 5; }
        else if (s.compare(i, 4, "null") == 0) { i += 4; }
        else {
            char *end = nullptr;
            v.num = std::strtod(s.c_str() + i, &end);
            if (end == s.c_str() + i) fail("unexpected character");
            v.type = Value::Number;
            i = end - s.c_str();
        }
        return v;
    }
    std::string string()
    {
        std::string out;
        i++;   // opening quote
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char e = s[++i];
                out += e == 'n' ? '\n' : e == 't' ? '\t' : e;
            } else out += s[i];
            i++;
        }
        if (i >= s.size()) fail("unterminated string");
        i++;
        return out;
    }
};

inline Value read_json(const std::string &file)
{
    std::ifstream f(file);
    if (!f) throw std::runtime_error(file + ": cannot open file");     // json_parser_error's wording
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    return Parser(text).parse();
}

}  // namespace json
}  // namespace visgeom_b200
