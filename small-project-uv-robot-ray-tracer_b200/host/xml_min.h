// xml_min.h -- a small XML element reader/writer, enough for the route files
// (positions/*.xml).  Output formatting matches what the reference's tinyxml2 printer
// produces for these documents: 4-space indentation, "%.8g" floats, LF line ends.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace uvrt_xml {

struct Element {
    std::string name, text;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<Element>> children;

    Element* child(const std::string& n) const
    {
        for (auto& c : children) if (c->name == n) return c.get();
        return nullptr;
    }
    const std::string* attr(const std::string& n) const
    {
        for (auto& a : attrs) if (a.first == n) return &a.second;
        return nullptr;
    }
    Element* add(const std::string& n)
    {
        children.emplace_back(new Element());
        children.back()->name = n;
        return children.back().get();
    }
    bool int_text(int* out) const
    {
        char* e = nullptr;
        long v = strtol(text.c_str(), &e, 10);
        if (e == text.c_str()) return false;
        *out = (int)v;
        return true;
    }
    bool float_text(float* out) const
    {
        char* e = nullptr;
        float v = strtof(text.c_str(), &e);
        if (e == text.c_str()) return false;
        *out = v;
        return true;
    }
    bool float_attr(const std::string& n, float* out) const
    {
        const std::string* s = attr(n);
        if (!s) return false;
        char* e = nullptr;
        float v = strtof(s->c_str(), &e);
        if (e == s->c_str()) return false;
        *out = v;
        return true;
    }
};

inline std::string fmt_float(float v)
{
    char b[64];
    snprintf(b, sizeof b, "%.8g", v);
    return b;
}

class Reader {
public:
    explicit Reader(const std::string& s) : p(s.c_str()), end(s.c_str() + s.size()) {}
    // returns the document's first element, or null on malformed input
    std::unique_ptr<Element> parse()
    {
        skip_misc();
        if (p >= end || *p != '<') return nullptr;
        auto e = element(0);
        return ok ? std::move(e) : nullptr;
    }

private:
    const char* p;
    const char* end;
    bool ok = true;

    void ws() { while (p < end && isspace((unsigned char)*p)) p++; }
    bool starts(const char* w) const { size_t n = strlen(w); return (size_t)(end - p) >= n && !memcmp(p, w, n); }
    void skip_until(const char* w)
    {
        size_t n = strlen(w);
        while (p < end && !starts(w)) p++;
        if (p < end) p += n; else ok = false;
    }
    void skip_misc()
    {
        for (;;) {
            ws();
            if (starts("<?")) skip_until("?>");
            else if (starts("<!--")) skip_until("-->");
            else if (starts("<!")) skip_until(">");
            else break;
            if (!ok) break;
        }
    }
    static std::string unescape(const std::string& s)
    {
        std::string o;
        for (size_t i = 0; i < s.size(); i++) {
            if (s[i] != '&') { o += s[i]; continue; }
            if (!s.compare(i, 5, "&amp;")) { o += '&'; i += 4; }
            else if (!s.compare(i, 4, "&lt;")) { o += '<'; i += 3; }
            else if (!s.compare(i, 4, "&gt;")) { o += '>'; i += 3; }
            else if (!s.compare(i, 6, "&quot;")) { o += '"'; i += 5; }
            else if (!s.compare(i, 6, "&apos;")) { o += '\''; i += 5; }
            else o += s[i];
        }
        return o;
    }
    std::string name_()
    {
        const char* b = p;
        while (p < end && !isspace((unsigned char)*p) && *p != '>' && *p != '/' && *p != '=') p++;
        return std::string(b, p);
    }
    std::unique_ptr<Element> element(int depth)
    {
        std::unique_ptr<Element> e(new Element());
        if (depth > 64) { ok = false; return e; }
        p++; // '<'
        e->name = name_();
        if (e->name.empty()) { ok = false; return e; }
        for (;;) {
            ws();
            if (p >= end) { ok = false; return e; }
            if (*p == '/') {
                if (p + 1 < end && p[1] == '>') { p += 2; return e; }
                ok = false;
                return e;
            }
            if (*p == '>') { p++; break; }
            std::string an = name_();
            ws();
            if (an.empty() || p >= end || *p != '=') { ok = false; return e; }
            p++;
            ws();
            if (p >= end || (*p != '"' && *p != '\'')) { ok = false; return e; }
            char q = *p++;
            const char* b = p;
            while (p < end && *p != q) p++;
            if (p >= end) { ok = false; return e; }
            e->attrs.emplace_back(an, unescape(std::string(b, p)));
            p++;
        }
        for (;;) {
            const char* b = p;
            while (p < end && *p != '<') p++;
            e->text += unescape(std::string(b, p));
            if (p >= end) { ok = false; return e; }
            if (starts("</")) {
                p += 2;
                std::string n = name_();
                ws();
                if (n != e->name || p >= end || *p != '>') { ok = false; return e; }
                p++;
                // trim
                size_t a = e->text.find_first_not_of(" \t\r\n");
                size_t z = e->text.find_last_not_of(" \t\r\n");
                e->text = a == std::string::npos ? "" : e->text.substr(a, z - a + 1);
                return e;
            }
            if (starts("<!--")) { skip_until("-->"); if (!ok) return e; continue; }
            if (starts("<?")) { skip_until("?>"); if (!ok) return e; continue; }
            e->children.push_back(element(depth + 1));
            if (!ok) return e;
        }
    }
};

inline void write(const Element& e, std::string& out, int depth)
{
    out.append((size_t)depth * 4, ' ');
    out += "<" + e.name;
    for (auto& a : e.attrs) out += " " + a.first + "=\"" + a.second + "\"";
    if (e.children.empty() && e.text.empty()) { out += "/>\n"; return; }
    out += ">";
    if (!e.children.empty()) {
        out += "\n";
        for (auto& c : e.children) write(*c, out, depth + 1);
        out.append((size_t)depth * 4, ' ');
    } else
        out += e.text;
    out += "</" + e.name + ">\n";
}

} // namespace uvrt_xml
