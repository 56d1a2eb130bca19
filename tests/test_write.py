"""Byte-level checks of the writer mirror (stochastic-inventory_b200/write.py) against Java's documented
`Double.toString` / `BigDecimal.setScale` behaviour (WriteToCsv.java:41-121)."""
import importlib

import pytest

import sdpb200 as S

w = importlib.import_module(S.package.__name__ + ".write")


def test_java_double_to_string():
    cases = {332.1462439016628: "332.1462439016628", 100.0: "100.0", 0.001: "0.001", 1e7: "1.0E7", 9999999.0: "9999999.0",
             1.5e-5: "1.5E-5", 1e22: "1.0E22", 0.5: "0.5", 12.5: "12.5", -67.0: "-67.0", 0.0: "0.0",
             1.0E-4: "1.0E-4", 123456789.125: "1.23456789125E8", -17.800000000000008: "-17.800000000000008",
             float("inf"): "Infinity", float("-inf"): "-Infinity", 4.9e-324: "4.9E-324", 1.7976931348623157e308: "1.7976931348623157E308"}
    for x, s in cases.items():
        assert w.java_double(x) == s, (x, w.java_double(x), s)
    assert w.java_double(float("nan")) == "NaN" and w.java_double(-0.0) == "-0.0"


def test_csv_writers_bytes(tmp_path):
    data = [[1.0, 2.5, -3.25], [100.0, 0.0, 12345678.0]]
    p = tmp_path / "a.csv"
    w.WriteToCsv().writeArrayCSV(data, str(p))
    assert p.read_bytes() == b"1.0,2.5,-3.25,\n100.0,0.0,1.2345678E7,\n"
    w.WriteToCsv().writeArrayCSVLabel(data, 5, -2, str(p))
    assert p.read_bytes() == b"x|R,-2,-1,0,\n5.0,1.0,2.5,-3.25,\n6.0,100.0,0.0,1.2345678E7,\n"
    w.WriteToCsv().writeArrayExcel([[1.0, 0.1], [332.1462439016628, 67.0]], str(p))
    assert p.read_bytes() == b"1.0\t0.1\t\n332.1462439016628\t67.0\t\n"
    w.WriteToExcelTxt().writeArrayToTxt([[2.0]], str(p))
    assert p.read_bytes() == b"2.0\t\n"
    w.WriteToCsv.writeToFile(str(p), "tail")
    assert p.read_bytes() == b"2.0\t\ntail\n"
    # new BigDecimal(0.1).setScale(2) needs rounding: the reference throws, so does the mirror
    with pytest.raises(ArithmeticError):
        w.WriteToCsv().writeArrayCSV([[0.1]], str(p))
