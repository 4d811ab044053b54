// json.h — minimal JSON reader/writer for the scene format and the benchmark JSON schema.
// (The reference uses Go's encoding/json; only what its schema needs is implemented here.)
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace gort {
namespace json {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;  // insertion order kept

    const Value* get(const std::string& key) const {
        if (kind != Object) return nullptr;
        // encoding/json keeps the LAST duplicate key
        const Value* found = nullptr;
        for (auto& kv : obj)
            if (kv.first == key) found = kv.second.get();
        return found;
    }
    bool is_number() const { return kind == Number; }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }
};

class Parser {
   public:
    Parser(const char* s, size_t n) : p_(s), end_(s + n) {}
    ValuePtr parse(std::string& err) {
        ValuePtr v = value(err);
        if (!v) return nullptr;
        ws();
        if (p_ != end_) {
            err = "trailing characters after JSON value";
            return nullptr;
        }
        return v;
    }

   private:
    static constexpr int kMaxDepth = 1000;
    const char* p_;
    const char* end_;
    int depth_ = 0;
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++;
    }
    ValuePtr value(std::string& err) {
        ws();
        if (p_ >= end_) {
            err = "unexpected end of JSON input";
            return nullptr;
        }
        char c = *p_;
        if (c == '{' || c == '[') {
            // The parser (and the tree's destructor) recurse once per nesting level, and scene text can arrive over the
            // network (chunk farm): bounded like Go's decoder (encoding/json stops at 10 000; a scene nests 4 deep)
            if (depth_ >= kMaxDepth) {
                err = "exceeded max depth";
                return nullptr;
            }
            depth_++;
            ValuePtr v = c == '{' ? object(err) : array(err);
            depth_--;
            return v;
        }
        if (c == '"') {
            auto v = std::make_shared<Value>();
            v->kind = Value::String;
            if (!string(v->str, err)) return nullptr;
            return v;
        }
        if (c == 't' || c == 'f' || c == 'n') return literal(err);
        return number(err);
    }
    ValuePtr literal(std::string& err) {
        auto v = std::make_shared<Value>();
        auto match = [&](const char* w) {
            size_t n = strlen(w);
            if ((size_t)(end_ - p_) >= n && strncmp(p_, w, n) == 0) {
                p_ += n;
                return true;
            }
            return false;
        };
        if (match("true")) {
            v->kind = Value::Bool;
            v->b = true;
        } else if (match("false")) {
            v->kind = Value::Bool;
            v->b = false;
        } else if (match("null")) {
            v->kind = Value::Null;
        } else {
            err = "invalid literal";
            return nullptr;
        }
        return v;
    }
    ValuePtr number(std::string& err) {
        const char* s = p_;
        if (p_ < end_ && (*p_ == '-' || *p_ == '+')) p_++;
        bool digits = false;
        while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
            if (*p_ >= '0' && *p_ <= '9') digits = true;
            p_++;
        }
        if (!digits) {
            err = "invalid character looking for beginning of value";
            return nullptr;
        }
        std::string tok(s, p_ - s);
        auto v = std::make_shared<Value>();
        v->kind = Value::Number;
        v->num = strtod(tok.c_str(), nullptr);
        return v;
    }
    bool string(std::string& out, std::string& err) {
        p_++;  // opening quote
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') {
                p_++;
                if (p_ >= end_) break;
                switch (*p_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (end_ - p_ < 5) { err = "bad \\u escape"; return false; }
                        unsigned cp = (unsigned)strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
                        p_ += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += *p_; break;
                }
                p_++;
            } else {
                out += *p_++;
            }
        }
        if (p_ >= end_) {
            err = "unterminated string";
            return false;
        }
        p_++;  // closing quote
        return true;
    }
    ValuePtr array(std::string& err) {
        auto v = std::make_shared<Value>();
        v->kind = Value::Array;
        p_++;
        ws();
        if (p_ < end_ && *p_ == ']') {
            p_++;
            return v;
        }
        for (;;) {
            ValuePtr e = value(err);
            if (!e) return nullptr;
            v->arr.push_back(e);
            ws();
            if (p_ < end_ && *p_ == ',') {
                p_++;
                continue;
            }
            if (p_ < end_ && *p_ == ']') {
                p_++;
                return v;
            }
            err = "expected ',' or ']' in array";
            return nullptr;
        }
    }
    ValuePtr object(std::string& err) {
        auto v = std::make_shared<Value>();
        v->kind = Value::Object;
        p_++;
        ws();
        if (p_ < end_ && *p_ == '}') {
            p_++;
            return v;
        }
        for (;;) {
            ws();
            if (p_ >= end_ || *p_ != '"') {
                err = "expected string key in object";
                return nullptr;
            }
            std::string key;
            if (!string(key, err)) return nullptr;
            ws();
            if (p_ >= end_ || *p_ != ':') {
                err = "expected ':' after object key";
                return nullptr;
            }
            p_++;
            ValuePtr e = value(err);
            if (!e) return nullptr;
            v->obj.emplace_back(key, e);
            ws();
            if (p_ < end_ && *p_ == ',') {
                p_++;
                continue;
            }
            if (p_ < end_ && *p_ == '}') {
                p_++;
                return v;
            }
            err = "expected ',' or '}' in object";
            return nullptr;
        }
    }
};

inline ValuePtr parse(const char* s, size_t n, std::string& err) {
    Parser p(s, n);
    return p.parse(err);
}

}  // namespace json
}  // namespace gort
