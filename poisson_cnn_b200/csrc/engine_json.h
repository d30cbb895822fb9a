// Minimal JSON reader for the model configs handed to pcnn_create (the reference's experiment JSONs:
// poisson_CNN/experiments/*.json sections hpnn_model / dbcnn_model).  Objects, arrays, strings, numbers, true/false/null.
#pragma once
#include <cctype>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pcnn {
namespace json {

struct Value {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    double num = 0.0;
    bool b = false;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value* find(const std::string& key) const {
        if (type != Obj) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const std::string& key) const { const Value* v = find(key); return v && v->type != Null; }
    const Value& at(const std::string& key) const {
        const Value* v = find(key);
        if (!v) throw std::runtime_error("config: missing key '" + key + "'");
        return *v;
    }
    double number(const std::string& key, double dflt) const {
        const Value* v = find(key);
        if (!v || v->type == Null) return dflt;
        if (v->type == Bool) return v->b ? 1.0 : 0.0;
        if (v->type != Num) throw std::runtime_error("config: '" + key + "' must be a number");
        return v->num;
    }
    bool boolean(const std::string& key, bool dflt) const {
        const Value* v = find(key);
        if (!v || v->type == Null) return dflt;
        if (v->type == Bool) return v->b;
        if (v->type == Num) return v->num != 0.0;
        throw std::runtime_error("config: '" + key + "' must be a boolean");
    }
    std::string string(const std::string& key, const std::string& dflt) const {
        const Value* v = find(key);
        if (!v || v->type == Null) return dflt;
        if (v->type != Str) throw std::runtime_error("config: '" + key + "' must be a string");
        return v->str;
    }
    std::vector<int> ints(const std::string& key) const {
        const Value& v = at(key);
        if (v.type != Arr) throw std::runtime_error("config: '" + key + "' must be a list");
        std::vector<int> out;
        for (const auto& e : v.arr) {
            if (e.type != Num) throw std::runtime_error("config: '" + key + "' must be a list of numbers");
            out.push_back((int)e.num);
        }
        return out;
    }
};

class Parser {
public:
    explicit Parser(const char* s) : p_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (*p_) fail("trailing characters");
        return v;
    }

private:
    const char* p_;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("config JSON: ") + what); }
    void ws() { while (*p_ && std::isspace((unsigned char)*p_)) ++p_; }
    bool lit(const char* w) {
        const char* q = p_;
        while (*w) if (*q++ != *w++) return false;
        p_ = q;
        return true;
    }
    Value value() {
        ws();
        Value v;
        if (*p_ == '{') {
            v.type = Value::Obj;
            ++p_; ws();
            if (*p_ == '}') { ++p_; return v; }
            for (;;) {
                ws();
                if (*p_ != '"') fail("expected a string key");
                std::string k = str();
                ws();
                if (*p_++ != ':') fail("expected ':'");
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; return v; }
                fail("expected ',' or '}'");
            }
        }
        if (*p_ == '[') {
            v.type = Value::Arr;
            ++p_; ws();
            if (*p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == ']') { ++p_; return v; }
                fail("expected ',' or ']'");
            }
        }
        if (*p_ == '"') { v.type = Value::Str; v.str = str(); return v; }
        if (lit("true")) { v.type = Value::Bool; v.b = true; return v; }
        if (lit("false")) { v.type = Value::Bool; v.b = false; return v; }
        if (lit("null")) return v;
        char* end = nullptr;
        v.num = std::strtod(p_, &end);
        if (end == p_) fail("unexpected character");
        v.type = Value::Num;
        p_ = end;
        return v;
    }
    std::string str() {
        std::string out;
        ++p_;
        while (*p_ && *p_ != '"') {
            if (*p_ == '\\') {
                ++p_;
                switch (*p_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'u': p_ += 4; out += '?'; break;      // names in these configs are ASCII
                    default: out += *p_;
                }
                if (*p_) ++p_;
            } else {
                out += *p_++;
            }
        }
        if (*p_ != '"') fail("unterminated string");
        ++p_;
        return out;
    }
};

inline Value parse(const char* s) { return Parser(s).parse(); }

}  // namespace json
}  // namespace pcnn
