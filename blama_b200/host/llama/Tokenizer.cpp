// Tokenizer.cpp -- see Tokenizer.hpp.
#include "Tokenizer.hpp"
#include "UnicodeTables.hpp"

#include <algorithm>
#include <array>
#include <queue>

namespace bl::llama {
namespace {

// ---- code points -----------------------------------------------------------------------------------------------------------
struct Cp { uint32_t cp; uint32_t off; };      // code point and its byte offset in the text

// UTF-8 -> code points; a byte that does not start a valid sequence stands alone as U+FFFD (its bytes stay in the word)
std::vector<Cp> decodeUtf8(std::string_view s) {
    std::vector<Cp> out;
    out.reserve(s.size() + 1);
    const auto* p = reinterpret_cast<const unsigned char*>(s.data());
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
        const unsigned char c = p[i];
        int len = 1; uint32_t cp = c;
        if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07; }
        else if (c >= 0x80) { out.push_back({0xFFFD, uint32_t(i)}); ++i; continue; }
        bool ok = i + size_t(len) <= n;
        for (int j = 1; ok && j < len; ++j) { ok = (p[i + size_t(j)] & 0xC0) == 0x80; cp = (cp << 6) | (p[i + size_t(j)] & 0x3F); }
        if (ok && len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) ok = false;
        if (ok && len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) ok = false;
        if (!ok) { out.push_back({0xFFFD, uint32_t(i)}); ++i; continue; }
        out.push_back({cp, uint32_t(i)});
        i += size_t(len);
    }
    return out;
}
void appendUtf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += char(cp);
    else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
    else { out += char(0xF0 | (cp >> 18)); out += char(0x80 | ((cp >> 12) & 0x3F)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
}

template <size_t N> bool inRanges(const unicode::CodepointRange (&r)[N], uint32_t cp) {
    size_t lo = 0, hi = N;
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (cp < r[mid].first) hi = mid;
        else if (cp > r[mid].last) lo = mid + 1;
        else return true;
    }
    return false;
}
bool isLetter(uint32_t cp) {
    if (cp < 0x80) return (cp >= 'A' && cp <= 'Z') || (cp >= 'a' && cp <= 'z');
    return inRanges(unicode::kLetterRanges, cp);
}
bool isNumber(uint32_t cp) {
    if (cp < 0x80) return cp >= '0' && cp <= '9';
    return inRanges(unicode::kNumberRanges, cp);
}
// \s: the White_Space property (the set llama.cpp's unicode.cpp uses)
bool isSpace(uint32_t cp) {
    return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
           cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
bool isNewline(uint32_t cp) { return cp == '\r' || cp == '\n'; }
bool isOther(uint32_t cp) { return !isSpace(cp) && !isLetter(cp) && !isNumber(cp); }      // [^\s\p{L}\p{N}]
uint32_t asciiLower(uint32_t cp) { return (cp >= 'A' && cp <= 'Z') ? cp + 32 : cp; }

// end of the whitespace alternatives shared by every pattern, at a whitespace character i:  (\s*[\r\n]+ only when newlineRule) |
// \s+(?!\S) | \s+
size_t matchSpace(const std::vector<Cp>& c, size_t i, bool newlineRule) {
    const size_t n = c.size();
    size_t e = i;
    while (e < n && isSpace(c[e].cp)) ++e;
    if (newlineRule) {
        for (size_t k = e; k > i; --k) if (isNewline(c[k - 1].cp)) return k;      // \s*[\r\n]+ : through the LAST newline of the run
    }
    if (e == n) return e;                 // \s+(?!\S): nothing follows
    if (e - i >= 2) return e - 1;         // ... or give back one character so that whitespace follows
    return e;                             // \s+
}

// one match of the llama3 / qwen2 pattern starting at i (maxDigits 3 / 1)
size_t matchLlama3(const std::vector<Cp>& c, size_t i, int maxDigits) {
    const size_t n = c.size();
    const uint32_t a = c[i].cp;
    if (a == '\'' && i + 1 < n) {         // (?i:'s|'t|'re|'ve|'m|'ll|'d)
        const uint32_t x = asciiLower(c[i + 1].cp), y = i + 2 < n ? asciiLower(c[i + 2].cp) : 0;
        if (x == 's' || x == 't' || x == 'm' || x == 'd') return i + 2;
        if ((x == 'r' && y == 'e') || (x == 'v' && y == 'e') || (x == 'l' && y == 'l')) return i + 3;
    }
    // [^\r\n\p{L}\p{N}]?\p{L}+
    if (!isNewline(a) && !isLetter(a) && !isNumber(a) && i + 1 < n && isLetter(c[i + 1].cp)) {
        size_t j = i + 2;
        while (j < n && isLetter(c[j].cp)) ++j;
        return j;
    }
    if (isLetter(a)) {
        size_t j = i + 1;
        while (j < n && isLetter(c[j].cp)) ++j;
        return j;
    }
    if (isNumber(a)) {                    // \p{N}{1,3} | \p{N}
        size_t j = i + 1;
        while (j < n && int(j - i) < maxDigits && isNumber(c[j].cp)) ++j;
        return j;
    }
    {                                     //  ?[^\s\p{L}\p{N}]+[\r\n]*
        const size_t k = (a == ' ' && i + 1 < n && isOther(c[i + 1].cp)) ? i + 1 : i;
        if (isOther(c[k].cp)) {
            size_t j = k + 1;
            while (j < n && isOther(c[j].cp)) ++j;
            while (j < n && isNewline(c[j].cp)) ++j;
            return j;
        }
    }
    if (isSpace(a)) return matchSpace(c, i, true);
    return i + 1;
}

// one match of the GPT-2 pattern starting at i
size_t matchGpt2(const std::vector<Cp>& c, size_t i) {
    const size_t n = c.size();
    const uint32_t a = c[i].cp;
    if (a == '\'' && i + 1 < n) {         // 's|'t|'re|'ve|'m|'ll|'d  (case-sensitive)
        const uint32_t x = c[i + 1].cp, y = i + 2 < n ? c[i + 2].cp : 0;
        if (x == 's' || x == 't' || x == 'm' || x == 'd') return i + 2;
        if ((x == 'r' && y == 'e') || (x == 'v' && y == 'e') || (x == 'l' && y == 'l')) return i + 3;
    }
    const size_t k = (a == ' ' && i + 1 < n) ? i + 1 : i;      // the optional leading space of the three word classes
    for (auto pred : {isLetter, isNumber, isOther}) {
        if (k != i && pred(c[k].cp)) { size_t j = k + 1; while (j < n && pred(c[j].cp)) ++j; return j; }
        if (pred(a)) { size_t j = i + 1; while (j < n && pred(c[j].cp)) ++j; return j; }
    }
    if (isSpace(a)) return matchSpace(c, i, false);
    return i + 1;
}

// GPT-2's bytes_to_unicode: printable Latin-1 bytes map to themselves, the other 68 to U+0100 + k in byte order
struct ByteMap {
    std::array<uint32_t, 256> toCp{};
    std::unordered_map<uint32_t, uint8_t> toByte;
    std::array<std::string, 256> utf8;
    ByteMap() {
        uint32_t next = 0;
        for (int b = 0; b < 256; ++b) {
            const bool printable = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
            toCp[size_t(b)] = printable ? uint32_t(b) : 256 + next++;
            toByte[toCp[size_t(b)]] = uint8_t(b);
            appendUtf8(utf8[size_t(b)], toCp[size_t(b)]);
        }
    }
};
const ByteMap& byteMap() { static const ByteMap m; return m; }

BpeTokenizer::Pre preOf(const std::string& name) {
    if (name == "llama-bpe" || name == "llama3" || name == "llama-v3") return BpeTokenizer::Pre::Llama3;
    if (name == "qwen2") return BpeTokenizer::Pre::Qwen2;
    return BpeTokenizer::Pre::Gpt2;
}

} // namespace

BpeTokenizer::BpeTokenizer(Config config) : m_cfg(std::move(config)), m_pre(preOf(m_cfg.pre)), m_ignoreMerges(m_pre == Pre::Llama3) {
    if (m_cfg.types.size() != m_cfg.tokens.size()) m_cfg.types.assign(m_cfg.tokens.size(), 1);
    m_byText.reserve(m_cfg.tokens.size());
    for (size_t i = 0; i < m_cfg.tokens.size(); ++i) m_byText.emplace(m_cfg.tokens[i], Token(i));      // first id wins for duplicate texts
    m_rank.reserve(m_cfg.merges.size());
    for (size_t r = 0; r < m_cfg.merges.size(); ++r) m_rank.emplace(m_cfg.merges[r], int32_t(r));
    for (size_t i = 0; i < m_cfg.tokens.size(); ++i) {
        const int32_t t = m_cfg.types[i];
        if ((t == 2 || t == 3 || t == 4) && !m_cfg.tokens[i].empty()) m_specials.push_back(Token(i));
    }
    std::stable_sort(m_specials.begin(), m_specials.end(), [&](Token a, Token b) { return m_cfg.tokens[size_t(a)].size() > m_cfg.tokens[size_t(b)].size(); });
}

std::vector<std::string_view> BpeTokenizer::split(std::string_view text) const {
    std::vector<std::string_view> words;
    const std::vector<Cp> c = decodeUtf8(text);
    const size_t n = c.size();
    for (size_t i = 0; i < n;) {
        size_t j = m_pre == Pre::Gpt2 ? matchGpt2(c, i) : matchLlama3(c, i, m_pre == Pre::Llama3 ? 3 : 1);
        if (j <= i) j = i + 1;
        const size_t b0 = c[i].off, b1 = j < n ? c[j].off : text.size();
        words.push_back(text.substr(b0, b1 - b0));
        i = j;
    }
    return words;
}

void BpeTokenizer::bpeWord(const std::string& word, std::vector<Token>& out) const {
    if (word.empty()) return;
    if (m_ignoreMerges) {
        const auto whole = m_byText.find(word);
        if (whole != m_byText.end()) { out.push_back(whole->second); return; }
    }
    // symbols = the UTF-8 characters of the byte-mapped word, doubly linked
    struct Sym { int prev, next; uint32_t off, len; };
    std::vector<Sym> sym;
    for (size_t i = 0; i < word.size();) {
        const unsigned char ch = static_cast<unsigned char>(word[i]);
        const uint32_t len = ch < 0x80 ? 1 : ch < 0xE0 ? 2 : ch < 0xF0 ? 3 : 4;
        sym.push_back({int(sym.size()) - 1, int(sym.size()) + 1, uint32_t(i), uint32_t(std::min<size_t>(len, word.size() - i))});
        i += len;
    }
    sym.back().next = -1;
    struct Bigram { int left, right; int32_t rank; uint32_t size; };
    auto worse = [](const Bigram& a, const Bigram& b) { return a.rank > b.rank || (a.rank == b.rank && a.left > b.left); };
    std::priority_queue<Bigram, std::vector<Bigram>, decltype(worse)> queue(worse);
    std::string key;
    auto consider = [&](int l, int r) {
        if (l < 0 || r < 0) return;
        key.assign(word, sym[size_t(l)].off, sym[size_t(l)].len);
        key += ' ';
        key.append(word, sym[size_t(r)].off, sym[size_t(r)].len);
        const auto it = m_rank.find(key);
        if (it == m_rank.end()) return;
        queue.push({l, r, it->second, sym[size_t(l)].len + sym[size_t(r)].len});
    };
    for (int i = 1; i < int(sym.size()); ++i) consider(i - 1, i);
    while (!queue.empty()) {
        const Bigram b = queue.top();
        queue.pop();
        Sym& L = sym[size_t(b.left)];
        Sym& R = sym[size_t(b.right)];
        if (L.len == 0 || R.len == 0 || L.len + R.len != b.size || L.next != b.right) continue;      // one side was merged away meanwhile
        L.len += R.len;
        R.len = 0;
        L.next = R.next;
        if (R.next >= 0) sym[size_t(R.next)].prev = b.left;
        consider(L.prev, b.left);
        consider(b.left, L.next);
    }
    for (int i = 0; i >= 0; i = sym[size_t(i)].next) {
        const Sym& s = sym[size_t(i)];
        if (s.len == 0) continue;
        key.assign(word, s.off, s.len);
        const auto it = m_byText.find(key);
        if (it != m_byText.end()) { out.push_back(it->second); continue; }
        for (char ch : key) {                     // unknown piece: its single bytes, where those exist as tokens
            const auto bt = m_byText.find(std::string(1, ch));
            if (bt != m_byText.end()) out.push_back(bt->second);
        }
    }
}

std::vector<Token> BpeTokenizer::tokenize(std::string_view text, bool addSpecial, bool parseSpecial) const {
    std::vector<Token> out;
    if (addSpecial && m_cfg.addBos && m_cfg.bos != Token_Invalid) out.push_back(m_cfg.bos);

    // special-token partition (llama.cpp tokenizer_st_partition): fragments are raw text or one special token
    struct Fragment { bool special; Token token; size_t off, len; };
    std::vector<Fragment> frags{{false, Token_Invalid, 0, text.size()}};
    for (Token sp : m_specials) {
        const int32_t type = m_cfg.types[size_t(sp)];
        if (!parseSpecial && (type == 3 || type == 2)) continue;      // CONTROL / UNKNOWN texts are ordinary text then
        const std::string& needle = m_cfg.tokens[size_t(sp)];
        std::vector<Fragment> next;
        next.reserve(frags.size());
        for (const Fragment& f : frags) {
            if (f.special) { next.push_back(f); continue; }
            size_t pos = f.off;
            const size_t end = f.off + f.len;
            for (;;) {
                const size_t hit = text.substr(0, end).find(needle, pos);
                if (hit == std::string_view::npos) break;
                if (hit > pos) next.push_back({false, Token_Invalid, pos, hit - pos});
                next.push_back({true, sp, hit, needle.size()});
                pos = hit + needle.size();
            }
            if (pos < end) next.push_back({false, Token_Invalid, pos, end - pos});
        }
        frags.swap(next);
    }

    const ByteMap& bm = byteMap();
    std::string word;
    for (const Fragment& f : frags) {
        if (f.special) { out.push_back(f.token); continue; }
        for (std::string_view w : split(text.substr(f.off, f.len))) {
            word.clear();
            for (unsigned char b : w) word += bm.utf8[b];
            bpeWord(word, out);
        }
    }
    if (addSpecial && m_cfg.addEos && m_cfg.eos != Token_Invalid) out.push_back(m_cfg.eos);
    return out;
}

std::string BpeTokenizer::tokenToPiece(Token token, bool special) const {
    if (token < 0 || size_t(token) >= m_cfg.tokens.size()) return {};
    const int32_t type = m_cfg.types[size_t(token)];
    const std::string& text = m_cfg.tokens[size_t(token)];
    const bool attrSpecial = type == 2 || type == 3 || type == 5;          // UNKNOWN | CONTROL | UNUSED
    if (!special && attrSpecial) return {};
    if (attrSpecial || type == 4) return text;                             // printed verbatim
    if (type != 1) return {};
    // NORMAL: code points back to bytes (llama_decode_text); a code point outside the byte alphabet is kept as it is
    const ByteMap& bm = byteMap();
    std::string out;
    for (const Cp& c : decodeUtf8(text)) {
        const auto it = bm.toByte.find(c.cp);
        if (it != bm.toByte.end()) out += char(it->second);
        else appendUtf8(out, c.cp);
    }
    return out;
}

} // namespace bl::llama
