// Json.hpp -- the small JSON reader / writer behind the HTTP contract (reference server/code/http/HttpServerMain.cpp:37-94,
// 255-288, which uses nlohmann::json 3.12 with its defaults).  What the wire format depends on, restated here:
//   * objects are std::map-ordered: dump() writes keys in byte-wise ascending order, compact, no spaces;
//   * a C++ float stored in a json is widened to double and printed as the SHORTEST decimal that parses back to that double,
//     laid out by nlohmann's rules (fixed notation for decimal exponents in (-4, 15], otherwise d.ddde[+-]XX with at least two
//     exponent digits, integral values get a trailing ".0"); get<float>() narrows the parsed double back, so a float survives
//     the wire bit for bit -- which is what keeps a /verify_completion verdict identical to the in-memory one;
//   * integers without '.', 'e' parse as (u)int64; get<int>() / get<uint32_t>() / get<float>() convert like static_cast;
//   * strings: the two-character escapes, \u00XX for other control characters, everything else passed through as UTF-8.
//     nlohmann throws on invalid UTF-8 while dumping (the reference then dies); token pieces of a
//     byte-level BPE vocabulary can be partial UTF-8 sequences, so this writer substitutes U+FFFD instead (documented deviation).
// Digits: nlohmann prints with Grisu2, which is shortest in > 99.9 % of the cases and otherwise one digit longer; std::to_chars is
// always shortest.  Both always parse back to the same double, so parity across the wire is unaffected.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <variant>
#include <vector>

namespace bl::json {

class Value;
using Array = std::vector<Value>;
using Object = std::map<std::string, Value, std::less<>>;

struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };
struct TypeError : std::runtime_error { using std::runtime_error::runtime_error; };

class Value {
public:
    Value() = default;
    Value(std::nullptr_t) {}
    Value(bool b) : m_v(b) {}
    Value(int64_t i) : m_v(i) {}
    Value(uint64_t u) : m_v(u) {}
    Value(int i) : m_v(int64_t(i)) {}
    Value(uint32_t u) : m_v(uint64_t(u)) {}
    Value(double d) : m_v(d) {}
    Value(float f) : m_v(double(f)) {}           // widened exactly, as nlohmann stores it
    Value(std::string s) : m_v(std::move(s)) {}
    Value(const char* s) : m_v(std::string(s)) {}
    Value(Array a) : m_v(std::move(a)) {}
    Value(Object o) : m_v(std::move(o)) {}

    bool isNull() const { return std::holds_alternative<std::monostate>(m_v); }
    bool isString() const { return std::holds_alternative<std::string>(m_v); }
    bool isArray() const { return std::holds_alternative<Array>(m_v); }
    bool isObject() const { return std::holds_alternative<Object>(m_v); }
    bool isNumber() const { return std::holds_alternative<int64_t>(m_v) || std::holds_alternative<uint64_t>(m_v) || std::holds_alternative<double>(m_v); }

    const std::string& str() const { if (auto* s = std::get_if<std::string>(&m_v)) return *s; throw TypeError("type must be string, but is " + typeName()); }
    std::string& str() { if (auto* s = std::get_if<std::string>(&m_v)) return *s; throw TypeError("type must be string, but is " + typeName()); }
    const Array& array() const { if (auto* a = std::get_if<Array>(&m_v)) return *a; throw TypeError("type must be array, but is " + typeName()); }
    const Object& object() const { if (auto* o = std::get_if<Object>(&m_v)) return *o; throw TypeError("type must be object, but is " + typeName()); }
    Object& object() { if (isNull()) m_v = Object{}; if (auto* o = std::get_if<Object>(&m_v)) return *o; throw TypeError("type must be object, but is " + typeName()); }
    Array& array() { if (isNull()) m_v = Array{}; if (auto* a = std::get_if<Array>(&m_v)) return *a; throw TypeError("type must be array, but is " + typeName()); }

    // number -> T like nlohmann's get<T>() (static_cast from whichever number type was parsed; bool converts too)
    template <class T> T num() const {
        if (auto* i = std::get_if<int64_t>(&m_v)) return static_cast<T>(*i);
        if (auto* u = std::get_if<uint64_t>(&m_v)) return static_cast<T>(*u);
        if (auto* d = std::get_if<double>(&m_v)) return static_cast<T>(*d);
        if (auto* b = std::get_if<bool>(&m_v)) return static_cast<T>(*b);
        throw TypeError("type must be number, but is " + typeName());
    }
    // object access: a missing key reads as null (nlohmann's non-const operator[] inserts a null)
    const Value& operator[](std::string_view key) const {
        static const Value null;
        if (auto* o = std::get_if<Object>(&m_v)) { auto it = o->find(key); return it == o->end() ? null : it->second; }
        if (isNull()) return null;
        throw TypeError("cannot use operator[] with a string argument with " + typeName());
    }
    const Value* find(std::string_view key) const {
        if (auto* o = std::get_if<Object>(&m_v)) { auto it = o->find(key); return it == o->end() ? nullptr : &it->second; }
        return nullptr;
    }
    Value& set(std::string key) { return object()[std::move(key)]; }
    size_t size() const {
        if (auto* a = std::get_if<Array>(&m_v)) return a->size();
        if (auto* o = std::get_if<Object>(&m_v)) return o->size();
        return isNull() ? 0 : 1;
    }
    std::string typeName() const {
        switch (m_v.index()) { case 0: return "null"; case 1: return "boolean"; case 2: case 3: case 4: return "number"; case 5: return "string"; case 6: return "array"; default: return "object"; }
    }

    std::string dump() const { std::string out; dumpTo(out); return out; }
    void dumpTo(std::string& out) const;

private:
    std::variant<std::monostate, bool, int64_t, uint64_t, double, std::string, Array, Object> m_v;
};

// ---- writer ------------------------------------------------------------------------------------------------------------
// shortest round-trip digits of a finite non-zero double laid out by nlohmann's format_buffer rules (min_exp -4, max_exp 15)
inline void appendDouble(std::string& out, double v) {
    if (!std::isfinite(v)) { out += "null"; return; }
    if (std::signbit(v)) { out += '-'; v = -v; }
    if (v == 0.0) { out += "0.0"; return; }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);     // d[.ddd]e[+-]XX, shortest
    std::string_view s(sci, size_t(r.ptr - sci));
    const size_t epos = s.find('e');
    std::string digits;
    for (char c : s.substr(0, epos)) if (c != '.') digits += c;
    int e10 = 0;
    std::from_chars(s.data() + epos + (s[epos + 1] == '+' ? 2 : 1), s.data() + s.size(), e10);
    const int k = int(digits.size());
    const int n = e10 + 1;                      // position of the decimal point relative to the first digit
    constexpr int min_exp = -4, max_exp = 15;
    if (k <= n && n <= max_exp) { out += digits; out.append(size_t(n - k), '0'); out += ".0"; return; }
    if (0 < n && n <= max_exp) { out.append(digits, 0, size_t(n)); out += '.'; out.append(digits, size_t(n), std::string::npos); return; }
    if (min_exp < n && n <= 0) { out += "0."; out.append(size_t(-n), '0'); out += digits; return; }
    out += digits[0];
    if (k > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
    out += 'e';
    int e = n - 1;
    out += e < 0 ? '-' : '+';
    if (e < 0) e = -e;
    if (e < 10) out += '0';
    out += std::to_string(e);
}

inline void appendString(std::string& out, std::string_view s) {
    out += '"';
    const auto* p = reinterpret_cast<const unsigned char*>(s.data());
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
        const unsigned char c = p[i];
        if (c < 0x80) {
            switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20 ) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; }
                else out += char(c);
            }
            ++i;
            continue;
        }
        // multi-byte sequence: validate (no overlongs, no surrogates, <= U+10FFFF); invalid bytes become U+FFFD
        int len = 0; uint32_t cp = 0;
        if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07; }
        bool ok = len != 0 && i + size_t(len) <= n;
        for (int j = 1; ok && j < len; ++j) { ok = (p[i + size_t(j)] & 0xC0) == 0x80; cp = (cp << 6) | (p[i + size_t(j)] & 0x3F); }
        if (ok && len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) ok = false;
        if (ok && len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) ok = false;
        if (ok) { out.append(reinterpret_cast<const char*>(p + i), size_t(len)); i += size_t(len); }
        else { out += "\xEF\xBF\xBD"; ++i; }
    }
    out += '"';
}

inline void Value::dumpTo(std::string& out) const {
    switch (m_v.index()) {
    case 0: out += "null"; break;
    case 1: out += std::get<bool>(m_v) ? "true" : "false"; break;
    case 2: out += std::to_string(std::get<int64_t>(m_v)); break;
    case 3: out += std::to_string(std::get<uint64_t>(m_v)); break;
    case 4: appendDouble(out, std::get<double>(m_v)); break;
    case 5: appendString(out, std::get<std::string>(m_v)); break;
    case 6: {
        out += '[';
        bool first = true;
        for (const auto& v : std::get<Array>(m_v)) { if (!first) out += ','; first = false; v.dumpTo(out); }
        out += ']';
        break;
    }
    default: {
        out += '{';
        bool first = true;
        for (const auto& [k, v] : std::get<Object>(m_v)) { if (!first) out += ','; first = false; appendString(out, k); out += ':'; v.dumpTo(out); }
        out += '}';
    }
    }
}

// ---- reader (RFC 8259; duplicate keys: the last one wins, as with nlohmann) -----------------------------------------------
class Parser {
public:
    explicit Parser(std::string_view text) : m_s(text) {}
    Value parseDocument() {
        Value v = parseValue(0);
        skipWs();
        if (m_i != m_s.size()) fail("unexpected trailing characters");
        return v;
    }

private:
    std::string_view m_s;
    size_t m_i = 0;
    static constexpr int kMaxDepth = 256;

    [[noreturn]] void fail(const std::string& what) const { throw ParseError("parse error at byte " + std::to_string(m_i + 1) + ": " + what); }
    void skipWs() { while (m_i < m_s.size() && (m_s[m_i] == ' ' || m_s[m_i] == '\t' || m_s[m_i] == '\n' || m_s[m_i] == '\r')) ++m_i; }
    bool eat(char c) { if (m_i < m_s.size() && m_s[m_i] == c) { ++m_i; return true; } return false; }
    void expectWord(std::string_view w) { if (m_s.substr(m_i, w.size()) != w) fail("invalid literal"); m_i += w.size(); }

    Value parseValue(int depth) {
        if (depth > kMaxDepth) fail("nesting too deep");
        skipWs();
        if (m_i >= m_s.size()) fail("unexpected end of input");
        const char c = m_s[m_i];
        if (c == '{') return parseObject(depth);
        if (c == '[') return parseArray(depth);
        if (c == '"') return Value(parseString());
        if (c == 't') { expectWord("true"); return Value(true); }
        if (c == 'f') { expectWord("false"); return Value(false); }
        if (c == 'n') { expectWord("null"); return Value(); }
        if (c == '-' || (c >= '0' && c <= '9')) return parseNumber();
        fail("unexpected character");
    }
    Value parseObject(int depth) {
        ++m_i;
        Object o;
        skipWs();
        if (eat('}')) return Value(std::move(o));
        for (;;) {
            skipWs();
            if (m_i >= m_s.size() || m_s[m_i] != '"') fail("object key must be a string");
            std::string key = parseString();
            skipWs();
            if (!eat(':')) fail("expected ':'");
            o[std::move(key)] = parseValue(depth + 1);
            skipWs();
            if (eat(',')) continue;
            if (eat('}')) break;
            fail("expected ',' or '}'");
        }
        return Value(std::move(o));
    }
    Value parseArray(int depth) {
        ++m_i;
        Array a;
        skipWs();
        if (eat(']')) return Value(std::move(a));
        for (;;) {
            a.push_back(parseValue(depth + 1));
            skipWs();
            if (eat(',')) continue;
            if (eat(']')) break;
            fail("expected ',' or ']'");
        }
        return Value(std::move(a));
    }
    static void appendUtf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) out += char(cp);
        else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
        else { out += char(0xF0 | (cp >> 18)); out += char(0x80 | ((cp >> 12) & 0x3F)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
    }
    uint32_t parseHex4() {
        if (m_i + 4 > m_s.size()) fail("truncated \\u escape");
        uint32_t v = 0;
        for (int j = 0; j < 4; ++j) {
            const char c = m_s[m_i++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
            else fail("invalid \\u escape");
        }
        return v;
    }
    std::string parseString() {
        ++m_i;
        std::string out;
        for (;;) {
            if (m_i >= m_s.size()) fail("unterminated string");
            const unsigned char c = static_cast<unsigned char>(m_s[m_i++]);
            if (c == '"') break;
            if (c < 0x20) fail("control character in string");
            if (c != '\\') { out += char(c); continue; }
            if (m_i >= m_s.size()) fail("unterminated escape");
            const char e = m_s[m_i++];
            switch (e) {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u': {
                uint32_t cp = parseHex4();
                if (cp >= 0xD800 && cp <= 0xDBFF) {
                    if (m_i + 2 > m_s.size() || m_s[m_i] != '\\' || m_s[m_i + 1] != 'u') fail("lone high surrogate");
                    m_i += 2;
                    const uint32_t lo = parseHex4();
                    if (lo < 0xDC00 || lo > 0xDFFF) fail("invalid low surrogate");
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                } else if (cp >= 0xDC00 && cp <= 0xDFFF) fail("lone low surrogate");
                appendUtf8(out, cp);
                break;
            }
            default: fail("invalid escape");
            }
        }
        return out;
    }
    Value parseNumber() {
        const size_t start = m_i;
        bool isFloat = false;
        if (eat('-')) {}
        if (m_i >= m_s.size()) fail("invalid number");
        if (m_s[m_i] == '0') ++m_i;
        else if (m_s[m_i] >= '1' && m_s[m_i] <= '9') { while (m_i < m_s.size() && m_s[m_i] >= '0' && m_s[m_i] <= '9') ++m_i; }
        else fail("invalid number");
        if (m_i < m_s.size() && m_s[m_i] == '.') {
            isFloat = true; ++m_i;
            if (m_i >= m_s.size() || m_s[m_i] < '0' || m_s[m_i] > '9') fail("invalid number");
            while (m_i < m_s.size() && m_s[m_i] >= '0' && m_s[m_i] <= '9') ++m_i;
        }
        if (m_i < m_s.size() && (m_s[m_i] == 'e' || m_s[m_i] == 'E')) {
            isFloat = true; ++m_i;
            if (m_i < m_s.size() && (m_s[m_i] == '+' || m_s[m_i] == '-')) ++m_i;
            if (m_i >= m_s.size() || m_s[m_i] < '0' || m_s[m_i] > '9') fail("invalid number");
            while (m_i < m_s.size() && m_s[m_i] >= '0' && m_s[m_i] <= '9') ++m_i;
        }
        const char* b = m_s.data() + start; const char* e = m_s.data() + m_i;
        if (!isFloat) {        // integers that do not fit 64 bits fall through to double, as in nlohmann's lexer
            if (*b == '-') { int64_t v = 0; auto r = std::from_chars(b, e, v); if (r.ec == std::errc() && r.ptr == e) return Value(v); }
            else { uint64_t v = 0; auto r = std::from_chars(b, e, v); if (r.ec == std::errc() && r.ptr == e) return Value(v); }
        }
        double d = 0.0;
        auto r = std::from_chars(b, e, d);
        if (r.ec == std::errc::result_out_of_range) fail("number overflow");
        if (r.ec != std::errc() || r.ptr != e) fail("invalid number");
        return Value(d);
    }
};

inline Value parse(std::string_view text) { return Parser(text).parseDocument(); }

} // namespace bl::json
