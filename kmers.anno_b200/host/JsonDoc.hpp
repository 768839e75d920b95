// JsonDoc.hpp — a small order-preserving JSON document (parse, edit, serialise) for the GTO files the
// `genes` command rewrites (GeneCopyProcessor.java:165 `this.target.save(this.outputFile)`; Genome /
// Feature are external SEEDtk classes, so only the JSON they read and write is modelled here).
// Numbers keep their source text, object members keep their file order, strings are held decoded
// (UTF-8) and re-escaped on output.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace theseed {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object };
    Kind kind = Null;
    bool flag = false;
    std::string text;                                           // Number: source text; String: decoded value
    std::vector<JsonValue> items;                               // Array
    std::vector<std::pair<std::string, JsonValue>> members;     // Object, in file order

    static JsonValue str(std::string s) { JsonValue v; v.kind = String; v.text = std::move(s); return v; }
    static JsonValue array() { JsonValue v; v.kind = Array; return v; }

    JsonValue* find(const std::string& key) {
        if (kind != Object) return nullptr;
        for (auto& m : members) if (m.first == key) return &m.second;
        return nullptr;
    }
    const JsonValue* find(const std::string& key) const { return const_cast<JsonValue*>(this)->find(key); }
    /** string member or "" (also for null / missing / non-string) */
    std::string getString(const std::string& key) const {
        const JsonValue* v = find(key);
        return v && v->kind == String ? v->text : std::string();
    }
    JsonValue& set(const std::string& key, JsonValue v) {
        if (JsonValue* old = find(key)) { *old = std::move(v); return *old; }
        members.emplace_back(key, std::move(v));
        return members.back().second;
    }

    static JsonValue parse(const std::string& s) {
        size_t p = 0;
        JsonValue v = parseValue(s, p);
        skipWs(s, p);
        if (p != s.size()) fail(p, "trailing characters");
        return v;
    }

    std::string dump() const { std::string out; write(out); return out; }

private:
    [[noreturn]] static void fail(size_t p, const std::string& m) {
        throw std::runtime_error("JSON error at byte " + std::to_string(p) + ": " + m);
    }
    static void skipWs(const std::string& s, size_t& p) {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++;
    }
    static unsigned hex4(const std::string& s, size_t& p) {
        if (p + 4 > s.size()) fail(p, "bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            char c = s[p++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail(p, "bad \\u escape");
        }
        return v;
    }
    static void utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    static std::string parseString(const std::string& s, size_t& p) {
        if (p >= s.size() || s[p] != '"') fail(p, "expected string");
        p++;
        std::string out;
        while (p < s.size() && s[p] != '"') {
            char c = s[p++];
            if (c != '\\') { out.push_back(c); continue; }
            if (p >= s.size()) fail(p, "bad escape");
            char e = s[p++];
            switch (e) {
                case 'n': out.push_back('\n'); break;
                case 't': out.push_back('\t'); break;
                case 'r': out.push_back('\r'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'u': {
                    unsigned cp = hex4(s, p);
                    if (cp >= 0xD800 && cp < 0xDC00 && p + 1 < s.size() && s[p] == '\\' && s[p + 1] == 'u') {
                        size_t save = p;
                        p += 2;
                        unsigned lo = hex4(s, p);
                        if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else p = save;
                    }
                    utf8(out, cp);
                    break;
                }
                default: out.push_back(e);   // \" \\ \/
            }
        }
        if (p >= s.size()) fail(p, "unterminated string");
        p++;
        return out;
    }
    static JsonValue parseValue(const std::string& s, size_t& p) {
        skipWs(s, p);
        if (p >= s.size()) fail(p, "unexpected end");
        JsonValue v;
        char c = s[p];
        if (c == '"') { v.kind = String; v.text = parseString(s, p); }
        else if (c == '{') {
            v.kind = Object; p++;
            skipWs(s, p);
            if (p < s.size() && s[p] == '}') { p++; return v; }
            for (;;) {
                skipWs(s, p);
                std::string key = parseString(s, p);
                skipWs(s, p);
                if (p >= s.size() || s[p] != ':') fail(p, "expected ':'");
                p++;
                v.members.emplace_back(std::move(key), parseValue(s, p));
                skipWs(s, p);
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                fail(p, "expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = Array; p++;
            skipWs(s, p);
            if (p < s.size() && s[p] == ']') { p++; return v; }
            for (;;) {
                v.items.push_back(parseValue(s, p));
                skipWs(s, p);
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                fail(p, "expected ',' or ']'");
            }
        } else if (s.compare(p, 4, "true") == 0) { v.kind = Bool; v.flag = true; p += 4; }
        else if (s.compare(p, 5, "false") == 0) { v.kind = Bool; v.flag = false; p += 5; }
        else if (s.compare(p, 4, "null") == 0) { v.kind = Null; p += 4; }
        else {
            size_t b = p;
            while (p < s.size() && (s[p] == '-' || s[p] == '+' || s[p] == '.' || s[p] == 'e' || s[p] == 'E' || (s[p] >= '0' && s[p] <= '9'))) p++;
            if (p == b) fail(p, "unexpected character");
            v.kind = Number; v.text = s.substr(b, p - b);
        }
        return v;
    }
    static void writeString(std::string& out, const std::string& s) {
        out.push_back('"');
        for (unsigned char c : s) {
            switch (c) {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\t': out += "\\t"; break;
                case '\r': out += "\\r"; break;
                case '\b': out += "\\b"; break;
                case '\f': out += "\\f"; break;
                default:
                    if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                    else out.push_back((char)c);
            }
        }
        out.push_back('"');
    }
    void write(std::string& out) const {
        switch (kind) {
            case Null: out += "null"; break;
            case Bool: out += flag ? "true" : "false"; break;
            case Number: out += text; break;
            case String: writeString(out, text); break;
            case Array:
                out.push_back('[');
                for (size_t i = 0; i < items.size(); i++) { if (i) out.push_back(','); items[i].write(out); }
                out.push_back(']');
                break;
            case Object:
                out.push_back('{');
                for (size_t i = 0; i < members.size(); i++) {
                    if (i) out.push_back(',');
                    writeString(out, members[i].first);
                    out.push_back(':');
                    members[i].second.write(out);
                }
                out.push_back('}');
                break;
        }
    }
};

}  // namespace theseed
