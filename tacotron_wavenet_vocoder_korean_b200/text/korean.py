# coding: utf-8
"""Jamo-free Korean front-end: text -> the reference's 80-symbol jamo ids.

Follows text/korean.py:11-25 (symbol inventory: PAD '_', EOS '~', 19 leads U+1100.., 21 vowels U+1161.., 27
tails U+11A8.., punctuation, space), :140-150 (`tokenize` = normalise, then decompose every Hangul syllable with
`jamo.hangul_to_jamo`, then EOS) and :153-168 (`normalize`).  The `jamo` package is replaced by the Unicode
arithmetic it implements.  Host string processing only; the reference's english/etc dictionaries
(text/ko_dictionary.py) are data files a deployment supplies through `set_dictionaries`.
"""
import re

PAD = '_'
EOS = '~'
PUNC = '!\'(),-.:;?'
SPACE = ' '

JAMO_LEADS = "".join(chr(c) for c in range(0x1100, 0x1113))
JAMO_VOWELS = "".join(chr(c) for c in range(0x1161, 0x1176))
JAMO_TAILS = "".join(chr(c) for c in range(0x11A8, 0x11C3))

VALID_CHARS = JAMO_LEADS + JAMO_VOWELS + JAMO_TAILS + PUNC + SPACE
ALL_SYMBOLS = PAD + EOS + VALID_CHARS

char_to_id = {c: i for i, c in enumerate(ALL_SYMBOLS)}
id_to_char = {i: c for i, c in enumerate(ALL_SYMBOLS)}

_english_dictionary = {}
_etc_dictionary = {}


def set_dictionaries(english=None, etc=None):
    """Install replacement tables with the role of text/ko_dictionary.py's english_dictionary / etc_dictionary."""
    if english is not None:
        _english_dictionary.clear()
        _english_dictionary.update(english)
    if etc is not None:
        _etc_dictionary.clear()
        _etc_dictionary.update(etc)


def hangul_to_jamo(text):
    """Decompose precomposed syllables U+AC00..U+D7A3 into conjoining jamo; other characters pass through."""
    for ch in text:
        code = ord(ch) - 0xAC00
        if 0 <= code < 11172:
            lead, rest = divmod(code, 588)
            vowel, tail = divmod(rest, 28)
            yield chr(0x1100 + lead)
            yield chr(0x1161 + vowel)
            if tail:
                yield chr(0x11A7 + tail)
        else:
            yield ch


def jamo_to_korean(text):
    """Recompose lead+vowel(+tail) runs into syllables (text/korean.py:52-82)."""
    out = []
    i = 0
    n = len(text)
    while i < n:
        c = text[i]
        if c in JAMO_LEADS and i + 1 < n and text[i + 1] in JAMO_VOWELS:
            lead = ord(c) - 0x1100
            vowel = ord(text[i + 1]) - 0x1161
            tail = 0
            i += 2
            if i < n and text[i] in JAMO_TAILS:
                tail = ord(text[i]) - 0x11A7
                i += 1
            out.append(chr(0xAC00 + lead * 588 + vowel * 28 + tail))
        else:
            out.append(c)
            i += 1
    return "".join(out)


_LETTER_NAMES = dict(zip("ABCDEFGHIJKLMNOPQRSTUVWXYZ",
                         "에이 비 씨 디 이 에프 지 에이치 아이 제이 케이 엘 엠 엔 오 피 큐 알 에스 티 유 브이 더블유 엑스 와이 지".split()))
_UNITS = [('%', '퍼센트'), ('cm', '센치미터'), ('mm', '밀리미터'), ('km', '킬로미터'), ('kg', '킬로그람'), ('m', '미터')]
_DIGIT = [''] + list("일이삼사오육칠팔구")
_DIGIT0 = ['영'] + list("일이삼사오육칠팔구")
_SMALL = ['', '십', '백', '천']
_BIG = ['', '만', '억', '조', '경', '해']
_COUNT_DIGIT = [''] + "한 두 세 네 다섯 여섯 일곱 여덟 아홉".split()
_COUNT_TENS = {'십': '열', '두십': '스물', '세십': '서른', '네십': '마흔', '다섯십': '쉰', '여섯십': '예순',
               '일곱십': '일흔', '여덟십': '여든', '아홉십': '아흔'}
_COUNTERS = "(시|명|가지|살|마리|포기|송이|수|톨|통|점|개|벌|척|채|다발|그루|자루|줄|켤레|그릇|잔|마디|상자|사람|곡|병|판)"
_NUMBER = r"([+-]?\d[\d,]*)[\.]?\d*"


def _replace_all(text, table):
    if table and any(k in text for k in table):
        pat = re.compile('|'.join(re.escape(k) for k in sorted(table, key=len, reverse=True)))
        return pat.sub(lambda m: table[m.group()], text)
    return text


def _read_integer(digits, native):
    """Group-of-four reading of a decimal digit string (sino-Korean, or native numerals for counters)."""
    words = _COUNT_DIGIT if native else _DIGIT
    size = len(digits)
    out, group = "", []
    for pos, ch in enumerate(digits, start=1):
        v = int(ch)
        if v:
            group.append(words[v])
            group.append(_SMALL[(size - pos) % 4])
        if (size - pos) % 4 == 0 and group:
            out += "".join(group) + _BIG[(size - pos) // 4]
            group = []
    return out


def number_to_korean(match, is_count=False):
    """text/korean.py:221-295 behaviour: '3,600' -> '삼천육백', '19가지' -> '열아홉가지', '-12.35' -> '마이너스 십이쩜 삼오'."""
    if is_count:
        num_str, unit = match.group(1), match.group(2)
    else:
        num_str, unit = match.group(), ""
    num_str = num_str.replace(',', '')
    sign = num_str[0] if num_str[0] in '+-' else ''
    body = num_str[len(sign):]
    if body.count('.') > 1:
        raise ValueError("wrong number format: %r" % num_str)
    whole, _, frac = body.partition('.')
    whole = whole.lstrip('0')
    if not whole and not frac.strip('0'):
        return "영" + unit
    kor = _read_integer(whole, is_count) if whole else ""
    if is_count:
        if kor.startswith("한") and len(kor) > 1:
            kor = kor[1:]
        kor = _replace_all(kor, _COUNT_TENS)
    elif kor.startswith("일") and len(kor) > 1:
        kor = kor[1:]
    if frac:
        kor += "쩜 " + "".join(_DIGIT0[int(c)] for c in frac)
    if sign == '+':
        kor = "플러스 " + kor
    elif sign == '-':
        kor = "마이너스 " + kor
    return kor + unit


def normalize(text):
    text = text.strip()
    text = re.sub(r'\(\d+일\)', '', text)
    text = re.sub(r'\([⺀-⿕々〇〡-〩〸-〻㐀-䶵一-鿃豈-龎]+\)', '', text)
    text = _replace_all(text, _etc_dictionary)
    text = re.sub("([A-Za-z]+)", lambda m: _english_dictionary.get(m.group(), m.group()), text)
    text = re.sub('[a-zA-Z]+', lambda m: "".join(_LETTER_NAMES[c] for c in m.group()) if m.group().isupper() else m.group(), text)
    text = re.sub("([`\"'＂“‘])(.+?)([`\"'＂”’])", lambda m: "'%s'" % m.group(2), text)
    for unit, kor in _UNITS[:5]:
        text = text.replace(unit, kor)
    text = text.replace(_UNITS[5][0], _UNITS[5][1])
    text = re.sub(_NUMBER + _COUNTERS, lambda m: number_to_korean(m, True), text)
    text = re.sub(_NUMBER, lambda m: number_to_korean(m, False), text)
    return text


def tokenize(text, as_id=False):
    tokens = list(hangul_to_jamo(normalize(text)))
    if as_id:
        return [char_to_id[t] for t in tokens] + [char_to_id[EOS]]
    return tokens + [EOS]
