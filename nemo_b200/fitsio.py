"""
Minimal FITS image I/O for the files the hot path exchanges with the rest of Nemo (the reference uses astropy.io.fits, which is
not a dependency here):

  * the filter cache ``filter_<label>#<tile>.fits``: a primary HDU holding the float64 filter ``[F, H, W]`` with the keywords
    ``SIGNORM``, ``RW<n>_GHZ`` and ``RW<n>`` (nemo/filters.py:767-794);
  * ``RMSMap_<label>#<tile>.fits`` (filters.py:760-762) and ``areaMask#<tile>.fits`` (maps.saveFITS, nemo/maps.py:2371-2411), with
    the tile's WCS keywords.  The reference writes these two tile-compressed (RICE_1 / PLIO_1); here they are written as plain
    image HDUs of the same type (float32 / uint8) -- what its readers look for is "the first HDU with data"
    (completeness._loadTile, nemo/completeness.py:813-826), which a plain primary HDU satisfies.

Only what FITS 4.0 needs for that: 2880-byte blocks, 80-character fixed-format header cards, big-endian data, BITPIX 8 / 16 / 32 /
64 / -32 / -64, BSCALE / BZERO on reading.  Files written here open unchanged in astropy (and vice versa for plain image HDUs).
"""

import os
import re

import numpy as np

BLOCK = 2880
_BITPIX_OF = {'uint8': 8, 'int16': 16, 'int32': 32, 'int64': 64, 'float32': -32, 'float64': -64}
_DTYPE_OF = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


def _formatValue(value):
    """Fixed-format value field (columns 11-30 for numbers and logicals, quoted from column 11 for strings)."""
    if isinstance(value, (bool, np.bool_)):
        return ("T" if value else "F").rjust(20)
    if isinstance(value, (int, np.integer)):
        return ("%d" % int(value)).rjust(20)
    if isinstance(value, (float, np.floating)):
        v = float(value)
        if v != v or v in (float('inf'), float('-inf')):
            raise ValueError("FITS header values must be finite (got %r)" % (value,))
        text = repr(v).upper()            # shortest string that reads back to the same double
        if 'E' in text:
            mant, exp = text.split('E')
            if '.' not in mant:
                mant += '.0'
            text = mant + 'E' + exp
        elif '.' not in text:
            text += '.0'
        return text.rjust(20)
    text = str(value).replace("'", "''")
    return ("'%s'" % text.ljust(8)).ljust(20)


def _card(key, value=None, comment=None):
    key = str(key).upper()
    if len(key) > 8 or not re.match(r'^[A-Z0-9_-]*$', key):
        raise ValueError("'%s' is not a FITS keyword (at most 8 characters of A-Z, 0-9, _ and -)" % key)
    if value is None:
        card = key.ljust(8) + (("  " + comment) if comment else "")
    else:
        card = key.ljust(8) + "= " + _formatValue(value)
        if comment:
            card += " / " + comment
    if len(card) > 80:
        if isinstance(value, str):
            raise ValueError("string value of %s does not fit one header card" % key)
        card = card[:80]
    return card.ljust(80)


def _headerItems(header):
    if header is None:
        return []
    items = header.items() if hasattr(header, 'items') else header
    return [(k, v) for k, v in items]


_STRUCTURAL = ('SIMPLE', 'BITPIX', 'NAXIS', 'EXTEND', 'END', 'BSCALE', 'BZERO', 'PCOUNT', 'GCOUNT', 'XTENSION')


def writeImage(path, data, header=None, overwrite=True):
    """Writes `data` (numpy array of up to 999 axes; uint8, int16/32/64, float32/64; bool -> uint8) as the primary HDU of a new
    FITS file, followed by the keywords of `header` (dict or list of (key, value) pairs; NAXISn and the other structural
    keywords in it are ignored -- they come from the array)."""
    arr = np.asarray(data)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    name = arr.dtype.name
    if name not in _BITPIX_OF:
        raise ValueError("cannot write a %s array to FITS (use uint8, int16, int32, int64, float32 or float64)" % name)
    if os.path.exists(path):
        if not overwrite:
            raise OSError("%s exists" % path)
        os.remove(path)
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', _BITPIX_OF[name], 'array data type'),
             _card('NAXIS', arr.ndim, 'number of array dimensions')]
    for i, n in enumerate(arr.shape[::-1], start=1):        # NAXIS1 is the fastest axis
        cards.append(_card('NAXIS%d' % i, int(n)))
    cards.append(_card('EXTEND', True))
    for key, value in _headerItems(header):
        k = str(key).upper()
        if k in _STRUCTURAL or re.match(r'^NAXIS\d+$', k):
            continue
        if k in ('COMMENT', 'HISTORY', ''):
            cards.append((k.ljust(8) + str(value))[:80].ljust(80))
        else:
            cards.append(_card(k, value))
    cards.append('END'.ljust(80))
    head = ''.join(cards).encode('ascii')
    head += b' ' * (-len(head) % BLOCK)
    body = np.ascontiguousarray(arr, dtype=_DTYPE_OF[_BITPIX_OF[name]]).tobytes()
    with open(path, 'wb') as fh:
        fh.write(head)
        fh.write(body)
        fh.write(b'\0' * (-len(body) % BLOCK))


def _parseValue(field):
    """Value of a card from the text after '= ' (comment included)."""
    text = field.strip()
    if text.startswith("'"):
        out, i = [], 1
        while i < len(text):
            if text[i] == "'":
                if i + 1 < len(text) and text[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(text[i])
            i += 1
        return ''.join(out).rstrip()
    text = text.split('/', 1)[0].strip()
    if text == 'T':
        return True
    if text == 'F':
        return False
    if text == '':
        return None
    try:
        return int(text)
    except ValueError:
        pass
    try:
        return float(text.upper().replace('D', 'E'))
    except ValueError:
        return text


def _readHeader(fh):
    """Header of the HDU at the file position: ({key: value} in file order, number of bytes read), or (None, 0) at end of file."""
    header, nRead = {}, 0
    while True:
        block = fh.read(BLOCK)
        if len(block) == 0 and nRead == 0:
            return None, 0
        if len(block) < BLOCK:
            raise OSError("truncated FITS header")
        nRead += BLOCK
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode('ascii', errors='replace')
            key = card[:8].strip()
            if key == 'END':
                return header, nRead
            if card[8:10] == '= ':
                header[key] = _parseValue(card[10:])
            elif key in ('COMMENT', 'HISTORY'):
                header.setdefault(key, [])
                header[key].append(card[8:].rstrip())


def _dataBytes(header):
    naxis = int(header.get('NAXIS', 0))
    if naxis == 0:
        return 0, ()
    shape = tuple(int(header['NAXIS%d' % i]) for i in range(naxis, 0, -1))
    n = abs(int(header['BITPIX'])) // 8 * int(header.get('GCOUNT', 1)) * (int(header.get('PCOUNT', 0)) + int(np.prod(shape)))
    return n, shape


def readImage(path, hdu=None, extName=None):
    """(data, header) of the first HDU that holds an image array, or of HDU number `hdu` (0 = primary), or of the extension
    whose EXTNAME is `extName` (completeness._loadTile: the tile's extension of a multi-extension file).  Tile-compressed
    images (binary tables with ZIMAGE = T) are not decoded: they raise."""
    with open(path, 'rb') as fh:
        index = 0
        while True:
            header, _ = _readHeader(fh)
            if header is None:
                raise OSError("no image data in %s" % path if extName is None else "no extension '%s' in %s" % (extName, path))
            nBytes, shape = _dataBytes(header)
            isImage = nBytes > 0 and header.get('XTENSION', 'IMAGE').strip() == 'IMAGE'
            if extName is not None:
                wanted = str(header.get('EXTNAME', '')).strip() == extName
            else:
                wanted = (hdu is None and nBytes > 0) or (hdu is not None and index == hdu)
            if wanted:
                if header.get('ZIMAGE') is True:
                    raise OSError("%s holds a tile-compressed image (ZCMPTYPE %s), which nemo_b200.fitsio does not decode" % (path, header.get('ZCMPTYPE')))
                if not isImage:
                    raise OSError("HDU %d of %s is not an image" % (index, path))
                raw = np.frombuffer(fh.read(nBytes), dtype=_DTYPE_OF[int(header['BITPIX'])]).reshape(shape)
                data = raw.astype(raw.dtype.newbyteorder('='))
                bscale, bzero = header.get('BSCALE', 1), header.get('BZERO', 0)
                if bscale != 1 or bzero != 0:
                    data = data * bscale + bzero
                return data, header
            fh.seek(nBytes + (-nBytes % BLOCK), 1)
            index += 1


def _imageCards(arr, first, header, extName=None):
    name = arr.dtype.name
    cards = [first, _card('BITPIX', _BITPIX_OF[name], 'array data type'), _card('NAXIS', arr.ndim, 'number of array dimensions')]
    for i, n in enumerate(arr.shape[::-1], start=1):
        cards.append(_card('NAXIS%d' % i, int(n)))
    if extName is not None:
        cards += [_card('PCOUNT', 0), _card('GCOUNT', 1), _card('EXTNAME', extName)]
    for key, value in _headerItems(header):
        k = str(key).upper()
        if k in _STRUCTURAL or k == 'EXTNAME' or re.match(r'^NAXIS\d+$', k):
            continue
        cards.append(_card(k, value))
    cards.append('END'.ljust(80))
    head = ''.join(cards).encode('ascii')
    return head + b' ' * (-len(head) % BLOCK)


def writeMEF(path, tiles, compressionType=None):
    """TileDict.saveMEF (nemo/maps.py:533-566): one IMAGE extension per tile, named after it, behind an empty primary HDU.
    tiles: list of (tileName, 2-d array, header dict or WCS object).  compressionType selects the pixel type only (see saveFITS)."""
    if os.path.exists(path):
        os.remove(path)
    with open(path, 'wb') as fh:
        head = ''.join([_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0), _card('EXTEND', True),
                        'END'.ljust(80)]).encode('ascii')
        fh.write(head + b' ' * (-len(head) % BLOCK))
        for tileName, data, header in tiles:
            arr = np.asarray(data)
            if compressionType is not None:
                arr = np.array(arr, dtype=np.uint8 if compressionType == 'PLIO_1' else np.float32)
            if arr.dtype == np.bool_:
                arr = arr.astype(np.uint8)
            if not hasattr(header, 'items'):
                header = wcsHeader(header)
            fh.write(_imageCards(arr, _card('XTENSION', 'IMAGE', 'Image extension'), header, extName=tileName))
            body = np.ascontiguousarray(arr, dtype=_DTYPE_OF[_BITPIX_OF[arr.dtype.name]]).tobytes()
            fh.write(body)
            fh.write(b'\0' * (-len(body) % BLOCK))


def wcsHeader(wcs):
    """The keywords of a WCS object (``.header``: astWCS.WCS and nemo_b200.wcs.CARWCS both carry one) as a plain dict."""
    if wcs is None or getattr(wcs, 'header', None) is None:
        return {}
    return {str(k): v for k, v in dict(wcs.header).items() if isinstance(v, (bool, int, float, str, np.integer, np.floating, np.bool_))}


def saveFITS(outputFileName, mapData, wcs, compressionType=None, version=None):
    """maps.saveFITS (nemo/maps.py:2371-2411): a map and its WCS keywords as a FITS image.  compressionType only selects the
    pixel type the reference would have compressed -- uint8 for 'PLIO_1' (masks), float32 for anything else -- the image itself
    is written uncompressed."""
    header = wcsHeader(wcs)
    header['NEMOVER'] = version if version is not None else 'nemo_b200'
    arr = np.asarray(mapData)
    if compressionType is not None:
        arr = np.array(arr, dtype=np.uint8 if compressionType == 'PLIO_1' else np.float32)
    writeImage(outputFileName, arr, header)


# ------------------------------------------------------------------------------------------------------
# Binary tables of float64 columns: the QFit.fits file of signals.fitQ (nemo/signals.py:1086-1117: one table per tile with columns
# Q, theta500Arcmin, z, as astropy's table_to_hdu writes them)
def writeTableMEF(path, tables):
    """tables: list of (extName, {column name: 1-d array}, header dict).  Every column is written as float64 (TFORM 'D')."""
    if os.path.exists(path):
        os.remove(path)
    with open(path, 'wb') as fh:
        head = ''.join([_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0), _card('EXTEND', True),
                        'END'.ljust(80)]).encode('ascii')
        fh.write(head + b' ' * (-len(head) % BLOCK))
        for extName, cols, header in tables:
            names = list(cols.keys())
            arrs = [np.asarray(cols[k], dtype=np.float64).ravel() for k in names]
            n = len(arrs[0]) if arrs else 0
            if any(len(a) != n for a in arrs):
                raise ValueError("columns of table %s differ in length" % extName)
            cards = [_card('XTENSION', 'BINTABLE', 'binary table extension'), _card('BITPIX', 8), _card('NAXIS', 2),
                     _card('NAXIS1', 8 * len(names), 'length of dimension 1'), _card('NAXIS2', n, 'length of dimension 2'),
                     _card('PCOUNT', 0), _card('GCOUNT', 1), _card('TFIELDS', len(names)), _card('EXTNAME', extName)]
            for i, k in enumerate(names, start=1):
                cards += [_card('TTYPE%d' % i, k), _card('TFORM%d' % i, 'D')]
            for key, value in _headerItems(header):
                cards.append(_card(key, value))
            cards.append('END'.ljust(80))
            head = ''.join(cards).encode('ascii')
            fh.write(head + b' ' * (-len(head) % BLOCK))
            body = np.ascontiguousarray(np.stack(arrs, axis=1) if names else np.zeros((0, 0)), dtype='>f8').tobytes()
            fh.write(body)
            fh.write(b'\0' * (-len(body) % BLOCK))


def readTable(path, extName):
    """({column: float64 array}, header) of the binary-table extension named `extName` (columns of TFORM 'D' only)."""
    with open(path, 'rb') as fh:
        while True:
            header, _ = _readHeader(fh)
            if header is None:
                raise OSError("no table '%s' in %s" % (extName, path))
            nBytes, _ = _dataBytes(header)
            if str(header.get('XTENSION', '')).strip() == 'BINTABLE' and str(header.get('EXTNAME', '')).strip() == extName:
                nCols, nRows = int(header['TFIELDS']), int(header['NAXIS2'])
                if any(str(header['TFORM%d' % i]).strip() not in ('D', '1D') for i in range(1, nCols + 1)):
                    raise OSError("table '%s' has columns that are not float64" % extName)
                raw = np.frombuffer(fh.read(8 * nCols * nRows), dtype='>f8').reshape(nRows, nCols)
                return {str(header['TTYPE%d' % i]).strip(): raw[:, i - 1].astype(np.float64) for i in range(1, nCols + 1)}, header
            fh.seek(nBytes + (-nBytes % BLOCK), 1)
