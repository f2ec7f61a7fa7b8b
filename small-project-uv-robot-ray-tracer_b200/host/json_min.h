// json_min.h -- a small recursive-descent JSON reader, enough for the JSON chunk of a .glb.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace uvrt_json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    bool has(const std::string& k) const { return kind == Object && obj.count(k) != 0; }
    const Value& operator[](const std::string& k) const
    {
        static const Value none;
        if (kind != Object) return none;
        auto it = obj.find(k);
        return it == obj.end() ? none : it->second;
    }
    const Value& operator[](size_t i) const
    {
        static const Value none;
        return (kind == Array && i < arr.size()) ? arr[i] : none;
    }
    size_t size() const { return kind == Array ? arr.size() : kind == Object ? obj.size() : 0; }
    // out-of-range and non-finite numbers (1e99, NaN) map to the default instead of an undefined conversion
    long long as_int(long long dflt = 0) const
    {
        return (kind == Number && num > -9.0e18 && num < 9.0e18) ? (long long)num : dflt;
    }
    bool is_number() const { return kind == Number; }
};

class Parser {
public:
    Parser(const char* p, size_t n) : s(p), end(p + n) {}
    bool parse(Value& out, std::string& err)
    {
        ok = true;
        out = value(0);
        ws();
        if (ok && s != end) fail("trailing characters");
        err = error;
        return ok;
    }

private:
    const char* s;
    const char* end;
    bool ok = true;
    std::string error;

    void fail(const char* m) { if (ok) { ok = false; error = m; } }
    void ws() { while (s < end && (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r')) s++; }
    bool lit(const char* w)
    {
        size_t n = strlen(w);
        if ((size_t)(end - s) >= n && !memcmp(s, w, n)) { s += n; return true; }
        return false;
    }
    std::string string_()
    {
        std::string out;
        s++; // opening quote
        while (s < end && *s != '"') {
            if (*s == '\\' && s + 1 < end) {
                s++;
                switch (*s) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u':
                    // keep \uXXXX escapes as '?': names are never used as keys we look up
                    out += '?';
                    s += (end - s > 4) ? 4 : (end - s - 1);
                    break;
                default: out += *s;
                }
                s++;
            } else
                out += *s++;
        }
        if (s >= end) fail("unterminated string"); else s++;
        return out;
    }
    Value value(int depth)
    {
        Value v;
        if (depth > 64) { fail("nesting too deep"); return v; }
        ws();
        if (s >= end) { fail("unexpected end"); return v; }
        if (*s == '{') {
            v.kind = Value::Object;
            s++;
            ws();
            if (s < end && *s == '}') { s++; return v; }
            while (ok) {
                ws();
                if (s >= end || *s != '"') { fail("expected key"); break; }
                std::string k = string_();
                ws();
                if (s >= end || *s != ':') { fail("expected ':'"); break; }
                s++;
                v.obj[k] = value(depth + 1);
                ws();
                if (s < end && *s == ',') { s++; continue; }
                if (s < end && *s == '}') { s++; break; }
                fail("expected ',' or '}'");
            }
        } else if (*s == '[') {
            v.kind = Value::Array;
            s++;
            ws();
            if (s < end && *s == ']') { s++; return v; }
            while (ok) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (s < end && *s == ',') { s++; continue; }
                if (s < end && *s == ']') { s++; break; }
                fail("expected ',' or ']'");
            }
        } else if (*s == '"') {
            v.kind = Value::String;
            v.str = string_();
        } else if (lit("true")) { v.kind = Value::Bool; v.b = true; }
        else if (lit("false")) { v.kind = Value::Bool; v.b = false; }
        else if (lit("null")) { v.kind = Value::Null; }
        else {
            std::string tmp;
            while (s < end && (isdigit((unsigned char)*s) || *s == '-' || *s == '+' || *s == '.' || *s == 'e' || *s == 'E')) tmp += *s++;
            if (tmp.empty()) { fail("unexpected character"); return v; }
            v.kind = Value::Number;
            v.num = strtod(tmp.c_str(), nullptr);
        }
        return v;
    }
};

} // namespace uvrt_json
