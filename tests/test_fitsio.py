"""
nemo_b200.fitsio (the reference's file formats without astropy): known-answer header cards as astropy writes them -- the layout
of examples/ACT-DR3-clusters/surveyMask.fits.gz in the reference tree, whose header the reader was checked against -- round trips
of every pixel type, the filter cache of nemo/filters.py:767-794 and the multi-extension area mask of nemo/maps.py:533-566.
"""

import os

import numpy as np
import pytest

from nemo_b200 import fitsio
from nemo_b200.wcs import CARWCS

# header of an astropy-written image as found in the reference's examples (fixed-format structural cards, free-format WCS cards)
ASTROPY_CARDS = [
    "SIMPLE  =                    T / conforms to FITS standard",
    "BITPIX  =                  -32 / array data type",
    "NAXIS   =                    2 / number of array dimensions",
    "NAXIS1  =                    3",
    "NAXIS2  =                    2",
    "CTYPE1  = 'RA---CEA'",
    "CTYPE2  = 'DEC--CEA'",
    "CUNIT1  = '        '",
    "CRVAL1  = 0",
    "CDELT1  = -0.008333333333333333",
    "CRPIX1  =    7281.600000000003",
    "CRPIX2  =               1260.0",
    "EQUINOX = 2000",
    "OBSERVER= 'O''Brien'           / quote inside a string",
    "SIGNORM =   1.234567890123E-05",
    "DVALUE  =              1.5D+03",
    "FLAG    =                    F",
    "END",
]


def _block(cards):
    text = ''.join(c.ljust(80) for c in cards)
    return (text + ' ' * (-len(text) % 2880)).encode('ascii')


def test_reader_parses_astropy_style_header(tmp_path):
    data = np.array([[1.5, -2.0, 3.25], [0.0, 1e-3, 7.0]], dtype='>f4')
    body = data.tobytes()
    path = str(tmp_path / 'a.fits')
    with open(path, 'wb') as fh:
        fh.write(_block(ASTROPY_CARDS) + body + b'\0' * (-len(body) % 2880))
    got, header = fitsio.readImage(path)
    assert got.dtype == np.float32 and got.dtype.isnative and np.array_equal(got, data.astype(np.float32))
    assert header['CTYPE1'] == 'RA---CEA' and header['CUNIT1'] == '' and header['CRVAL1'] == 0 and isinstance(header['CRVAL1'], int)
    assert header['CDELT1'] == -0.008333333333333333 and header['CRPIX1'] == 7281.600000000003 and header['CRPIX2'] == 1260.0
    assert header['OBSERVER'] == "O'Brien" and header['SIGNORM'] == 1.234567890123e-05 and header['DVALUE'] == 1500.0
    assert header['FLAG'] is False and header['SIMPLE'] is True and header['BITPIX'] == -32


def test_writer_emits_standard_fixed_format_cards(tmp_path):
    path = str(tmp_path / 'w.fits')
    arr = np.arange(6, dtype=np.float32).reshape(2, 3)
    fitsio.writeImage(path, arr, {'CTYPE1': 'RA---CAR', 'CRPIX1': 7281.600000000003, 'SIGNORM': 1.5e-05, 'RW1_GHZ': 148.0, 'N': 7,
                                  'NAXIS1': 99, 'OK': True})
    raw = open(path, 'rb').read()
    assert len(raw) == 2 * 2880
    cards = [raw[i:i + 80].decode('ascii').rstrip() for i in range(0, 2880, 80)]
    assert cards[:6] == ["SIMPLE  =                    T / conforms to FITS standard", "BITPIX  =                  -32 / array data type",
                         "NAXIS   =                    2 / number of array dimensions", "NAXIS1  =                    3",
                         "NAXIS2  =                    2", "EXTEND  =                    T"]
    assert "CTYPE1  = 'RA---CAR'" in cards and "CRPIX1  =    7281.600000000003" in cards and "SIGNORM =              1.5E-05" in cards
    assert "RW1_GHZ =                148.0" in cards and "N       =                    7" in cards and "OK      =                    T" in cards
    assert cards[cards.index("OK      =                    T") + 1] == "END"
    assert not any(c.startswith("NAXIS1  =                   99") for c in cards)       # structural keywords come from the array
    assert np.array_equal(np.frombuffer(raw[2880:2880 + 24], dtype='>f4'), arr.ravel())
    with pytest.raises(ValueError):
        fitsio.writeImage(path, arr, {'TOOLONGKEY': 1})
    with pytest.raises(ValueError):
        fitsio.writeImage(path, arr, {'BAD': float('nan')})


@pytest.mark.parametrize("dtype", ['uint8', 'int16', 'int32', 'int64', 'float32', 'float64', 'bool'])
def test_round_trip_of_every_pixel_type(tmp_path, dtype):
    rng = np.random.default_rng(1)
    arr = (rng.random((5, 7, 11)) * 200 - (0 if dtype in ('uint8', 'bool') else 100))
    arr = (arr > 100) if dtype == 'bool' else arr.astype(dtype)
    path = str(tmp_path / 'r.fits')
    fitsio.writeImage(path, arr, {'A': 1})
    got, header = fitsio.readImage(path)
    assert got.shape == arr.shape and np.array_equal(got, arr.astype(np.uint8) if dtype == 'bool' else arr)
    assert header['NAXIS'] == 3 and header['NAXIS1'] == 11 and header['NAXIS3'] == 5 and os.path.getsize(path) % 2880 == 0


def test_filter_cache_layout_round_trip(tmp_path):
    """filter_<label>#<tile>.fits of the reference: float64 [F, H, W] + SIGNORM, RW<n>_GHZ, RW<n> (filters.py:767-794), header
    values back bit for bit."""
    rng = np.random.default_rng(2)
    filt = rng.standard_normal((2, 12, 16))
    header = {'SIGNORM': 1.0 / 3.0, 'RW1_GHZ': 149.6, 'RW1': 0.6180339887498949, 'RW2_GHZ': 97.8, 'RW2': 1 - 0.6180339887498949}
    path = str(tmp_path / 'filter_Arnaud_M2e14_z0p4#1_2_3.fits')
    fitsio.writeImage(path, filt, header)
    got, h = fitsio.readImage(path)
    assert got.dtype == np.float64 and np.array_equal(got, filt)
    for k, v in header.items():
        assert h[k] == v


def test_save_fits_and_mef_carry_the_wcs(tmp_path):
    wcs = CARWCS((20, 30), 0.5 / 60, crval=(35.0, -42.5))
    rms = np.random.default_rng(3).random((20, 30))
    p1 = str(tmp_path / 'RMSMap_x#t.fits')
    fitsio.saveFITS(p1, rms, wcs, compressionType='RICE_1')
    got, h = fitsio.readImage(p1)
    assert got.dtype == np.float32 and np.array_equal(got, rms.astype(np.float32))
    assert h['CTYPE1'] == 'RA---CAR' and h['CRVAL2'] == -42.5 and h['CRPIX1'] == 16.0 and h['CDELT1'] == -0.5 / 60 and 'NEMOVER' in h
    masks = [('1_0_0', (rms > 0.5), wcs), ('1_0_1', (rms > 0.2), wcs)]
    p2 = str(tmp_path / 'areaMask.fits')
    fitsio.writeMEF(p2, masks, compressionType='PLIO_1')
    for name, m, _ in masks:
        got, h = fitsio.readImage(p2, extName=name)
        assert got.dtype == np.uint8 and np.array_equal(got, m.astype(np.uint8)) and h['EXTNAME'] == name and h['XTENSION'] == 'IMAGE'
    first, _ = fitsio.readImage(p2)           # "first HDU with data", what completeness._loadTile falls back to
    assert np.array_equal(first, masks[0][1].astype(np.uint8))
    with pytest.raises(OSError):
        fitsio.readImage(p2, extName='nope')


def test_tile_compressed_images_are_refused(tmp_path):
    cards = ["XTENSION= 'BINTABLE'", "BITPIX  =                    8", "NAXIS   =                    2", "NAXIS1  =                    8",
             "NAXIS2  =                    2", "PCOUNT  =                    0", "GCOUNT  =                    1",
             "ZIMAGE  =                    T", "ZCMPTYPE= 'RICE_1  '", "END"]
    primary = ["SIMPLE  =                    T", "BITPIX  =                    8", "NAXIS   =                    0", "END"]
    path = str(tmp_path / 'c.fits')
    with open(path, 'wb') as fh:
        fh.write(_block(primary) + _block(cards) + b'\0' * 2880)
    with pytest.raises(OSError, match='tile-compressed'):
        fitsio.readImage(path)
