# coding: utf-8
"""Generates tests/golden/ref_text.json by running THE REFERENCE'S OWN text/korean.py (normalize, number / English / quote
normalisation with its dictionaries, tokenize) on the 160 transcripts it ships (datasets/{moon,son}/*-recognition-All.json) and
on a set of sentences exercising numbers, units, counters, English letters and quotes.  The third-party `jamo` package
(hangul_to_jamo, h2j, j2h) is not installable here and is restated by Unicode arithmetic; everything else is the reference's.

    python tests/golden/make_reference_text_golden.py        (build container only)
"""
import importlib
import json
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
warnings.simplefilter('ignore')


def hangul_to_jamo(s):
    for ch in s:
        c = ord(ch)
        if 0xAC00 <= c <= 0xD7A3:
            i = c - 0xAC00
            yield chr(0x1100 + i // 588)
            yield chr(0x1161 + (i % 588) // 28)
            if i % 28:
                yield chr(0x11A7 + i % 28)
        else:
            yield ch


def main():
    jm = types.ModuleType('jamo')
    jm.hangul_to_jamo = hangul_to_jamo
    jm.h2j = lambda s: ''.join(hangul_to_jamo(s))
    jm.j2h = lambda lead, vowel, tail=None: chr(0xAC00 + (ord(lead) - 0x1100) * 588 + (ord(vowel) - 0x1161) * 28 + ((ord(tail) - 0x11A7) if tail else 0))
    sys.modules['jamo'] = jm
    pkg = types.ModuleType('reftext')
    pkg.__path__ = [os.path.join(REF, 'text')]
    sys.modules['reftext'] = pkg
    rk = importlib.import_module('reftext.korean')
    sents = []
    for spk in ('moon', 'son'):
        d = json.load(open(os.path.join(REF, 'datasets', spk, '%s-recognition-All.json' % spk), encoding='utf-8'))
        sents += [d[k] for k in sorted(d)]
    sents += ['오늘은 3일입니다.', '사과 2개와 배 10개', '1,234원입니다', '온도는 -3.5도', "그는 '안녕'이라고 말했다", '제 전화번호는 010-1234-5678 입니다.',
              '100%', '3시 15분', 'TV를 봤다', '12월 25일(화)', '2018년', '스물 한 살', '0.5초', 'KBS 뉴스 9', '30대 남성 2명', '1m 20cm', '5kg',
              '제1회 대회', '3.14', '1+1=2', '천 원짜리 3장', '7번째', 'A, B 그리고 C', '"따옴표" 테스트', '99마리의 양', '20,000명', '1/2', '2~3일',
              '안녕하세요?', '정말! 대단해...']
    out = []
    eng, etc = dict(rk.english_dictionary), dict(rk.etc_dictionary)
    for s in sents:
        try:
            c = dict(text=s, normalized=rk.normalize(s), ids=[int(i) for i in rk.tokenize(s, as_id=True)])
            # the same with the replacement tables of text/ko_dictionary.py emptied: this repository ships the algorithm, not that
            # data file (text.korean.set_dictionaries installs one)
            rk.english_dictionary.clear()
            rk.etc_dictionary.clear()
            try:
                c['ids_without_dictionaries'] = [int(i) for i in rk.tokenize(s, as_id=True)]
            finally:
                rk.english_dictionary.update(eng)
                rk.etc_dictionary.update(etc)
            out.append(c)
        except Exception as e:      # inputs the reference itself cannot tokenise (e.g. lower-case Latin letters) are recorded as such
            out.append(dict(text=s, error=type(e).__name__))
    sym = dict(all_symbols=rk.ALL_SYMBOLS, pad=rk.PAD, eos=rk.EOS)
    json.dump(dict(symbols=sym, cases=out), open(os.path.join(HERE, 'ref_text.json'), 'w', encoding='utf-8'), ensure_ascii=False, indent=0)
    print(len(out), 'sentences,', sum('error' in c for c in out), 'rejected by the reference itself')


if __name__ == '__main__':
    main()
