"""Host-side mirror of the reference's table writers, so experiment outputs
made from the GPU tables are byte-identical to what the Java drivers write:

    WriteToCsv         src/sdp/write/WriteToCsv.java:21-123   (writeToFile, writeArrayCSV, writeArrayCSVLabel,
                                                                writeArrayExcel)
    WriteToExcelTxt    src/sdp/write/WriteToExcelTxt.java:21-70 (writeToFile, writeArrayToTxt, writeArrayToExcel)

Java prints a double with `Double.toString` (shortest digits that round-trip -- JDK >= 19, the reference
builds with 21 -- in plain notation for 1e-3 <= |x| < 1e7 and `d.dddE[-]n` otherwise); `java_double` below
produces exactly that text.  `writeArrayCSV` goes through `new BigDecimal(x).setScale(2)`, which THROWS when
x has more than two exact binary decimals (`ArithmeticException: Rounding necessary`); the mirror raises the
same way instead of inventing a rounding the reference does not have.
"""
from __future__ import annotations

import math
from decimal import Decimal
from fractions import Fraction


def java_double(x: float) -> str:
    """`Double.toString(x)`."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    digits, exp = _shortest_digits(abs(x))  # value = 0.d1d2... x 10^exp
    if 1e-3 <= abs(x) < 1e7:
        if exp <= 0:
            body = "0." + "0" * (-exp) + digits
        elif exp >= len(digits):
            body = digits + "0" * (exp - len(digits)) + ".0"
        else:
            body = digits[:exp] + "." + digits[exp:]
        return sign + body
    mant = digits[0] + "." + (digits[1:] if len(digits) > 1 else "0")
    return f"{sign}{mant}E{exp - 1}"


def _shortest_digits(a: float):
    """Shortest decimal digit string that round-trips, and the decimal exponent (repr() is shortest too)."""
    r = repr(a)
    if "e" in r or "E" in r:
        m, e = r.lower().split("e")
        e = int(e)
    else:
        m, e = r, 0
    if "." in m:
        ip, fp = m.split(".")
    else:
        ip, fp = m, ""
    digits = (ip + fp).lstrip("0")
    lead = len((ip + fp)) - len((ip + fp).lstrip("0"))
    exp = len(ip) - lead + e
    digits = digits.rstrip("0") or "0"
    if len(digits) == 1:
        # Java prints at least two digits and, when one would do, takes the two-digit decimal CLOSEST to the
        # exact value among those that round-trip (Double.MIN_VALUE is "4.9E-324", not "5.0E-324")
        t = Decimal(a).scaleb(-(exp - 1)).quantize(Decimal("0.1"))  # d.d
        cand = str(t).replace(".", "")
        if len(cand) == 2 and float(f"0.{cand}e{exp}") == a:
            digits = cand.rstrip("0") or cand[0]
    return digits, exp


def _set_scale_2(x: float) -> float:
    """`new BigDecimal(x).setScale(2).doubleValue()`: exact, or ArithmeticException."""
    if math.isnan(x) or math.isinf(x):
        raise ArithmeticError("NumberFormatException: Infinite or NaN")
    f = Fraction(float(x)) * 100
    if f.denominator != 1:
        raise ArithmeticError("Rounding necessary")
    return float(Decimal(f.numerator) / Decimal(100))


class WriteToCsv:
    @staticmethod
    def writeToFile(fileName, s):
        """Append one line (WriteToCsv.java:21-33)."""
        with open(fileName, "a", encoding="utf-8", newline="") as f:
            f.write(str(s) + "\n")

    def writeArrayCSV(self, data, fileName):
        """WriteToCsv.java:41-58: every value as setScale(2).doubleValue(), trailing comma on every row."""
        with open(fileName, "w", encoding="utf-8", newline="") as f:
            for row in data:
                for v in row:
                    f.write(java_double(_set_scale_2(v)) + ",")
                f.write("\n")

    def writeArrayCSVLabel(self, data, minCash, minInventory, fileName):
        """WriteToCsv.java:66-98: first row `x|R`, inventory labels as integers; first column cash labels."""
        with open(fileName, "w", encoding="utf-8", newline="") as f:
            ncol = len(data[0])
            f.write("x|R,")
            for j in range(ncol):
                f.write(str(int(minInventory) + j) + ",")
            f.write("\n")
            for i, row in enumerate(data):
                f.write(java_double(_set_scale_2(float(int(minCash) + i))) + ",")
                for v in row:
                    f.write(java_double(_set_scale_2(v)) + ",")
                f.write("\n")

    def writeArrayExcel(self, data, fileName):
        """WriteToCsv.java:106-121: tab-separated Double.toString, trailing tab on every row."""
        with open(fileName, "w", encoding="utf-8", newline="") as f:
            for row in data:
                for v in row:
                    f.write(java_double(v) + "\t")
                f.write("\n")


class WriteToExcelTxt:
    writeToFile = staticmethod(WriteToCsv.writeToFile)

    def writeArrayToTxt(self, data, fileName):
        """WriteToExcelTxt.java:34-48."""
        WriteToCsv().writeArrayExcel(data, fileName)

    writeArrayToExcel = writeArrayToTxt  # WriteToExcelTxt.java:51-65: same bytes
